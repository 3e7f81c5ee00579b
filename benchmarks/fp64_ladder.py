"""Solo duration of the point-decompression kernel (k_decompress over 2^lg resident encodings) as a function of the share of its warps whose
2^252-3 ladder runs on the FP64 pipe (csrc/fd.cuh).  Prints the implied instruction rates of the two pipes.
Usage: python benchmarks/fp64_ladder.py [lg points, default 20] [alternative libbpmsm build]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import bulletproofs_b200 as bp
if len(sys.argv) > 2: bp.LIB_PATH = os.path.abspath(sys.argv[2])
import bench_msm

LG = int(sys.argv[1]) if len(sys.argv) > 1 else 20
n = 1 << LG
s1 = torch.cuda.Stream()
c1 = bp.Context(0, stream=s1.cuda_stream)
pts = c1.from_uniform_bytes(bench_msm.g_chain_uniform(n))
d_pts = torch.frombuffer(bytearray(pts), dtype=torch.uint8).cuda()
out = {"points": n, "lib": os.path.basename(bp.LIB_PATH), "solo": {}}
for share in (0, 2, 4, 6, 8):
    c1.set_fp64_share(share)
    bp.PointSet(c1, n=n, device_ptr=d_pts.data_ptr()).close(); c1.synchronize()     # one k_decompress launch over n points
    c1.prof_enable(True)
    for _ in range(3): bp.PointSet(c1, n=n, device_ptr=d_pts.data_ptr()).close()
    rep = c1.prof_report(); c1.prof_enable(False)
    ms = rep["k_decompress"][0] / rep["k_decompress"][1]
    out["solo"][share] = {"ms": round(ms, 4), "ns_per_point": round(1e6 * ms / n, 3), "Mpoints_per_s": round(n / ms / 1e3, 1)}
ms0, ms8 = out["solo"][0]["ms"], out["solo"][8]["ms"]
# thread-level instructions per second: the FP64 ladder is 249+5 squarings x 120 + 12 multiplications x 177 DFMA/DADD/DMUL; the integer one 254 x 44 + 12 x 72 IMAD.WIDE, plus ~36 multiplications outside the ladder either way
out["fp64_Tinstr_per_s_at_share_8"] = round(n * (254 * 120 + 12 * 177) / (ms8 * 1e-3) / 1e12, 2)
out["imad_wide_T_per_s_at_share_0"] = round(n * (254 * 44 + 48 * 72) / (ms0 * 1e-3) / 1e12, 2)
print(json.dumps(out))
