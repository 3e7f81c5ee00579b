"""Host cost of one device-path call of a reserved verifier (parameter upload + CUDA-graph launch) on an idle stream, against the same call
issued into a stream that already holds queued groups (driver back-pressure).  Prints one JSON object."""
import json, os, statistics, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import bulletproofs_b200 as bp
import bench

G, BATCH, N, M = 8, 1024, 64, 1
orc, og, proofs, Vs = bench.make_workload(BATCH, 0, M)
plen = len(proofs) // BATCH
stream = torch.cuda.Stream(); ctx = bp.Context(0, stream=stream.cuda_stream); gens = bp.Gens(ctx, N, M)
ver = bp.BatchVerifier(ctx, gens, bp.Transcript(bench.LABEL), N, M, BATCH, G)
d_p = torch.frombuffer(bytearray(proofs * G), dtype=torch.uint8).cuda(); d_v = torch.frombuffer(bytearray(Vs * G), dtype=torch.uint8).cuda()
d_out = torch.zeros(G * BATCH, dtype=torch.int32, device="cuda")
torch.cuda.synchronize()
idle, queued = [], []
for _ in range(30):
    t0 = time.perf_counter(); ver.run_device(d_p.data_ptr(), d_v.data_ptr(), d_out.data_ptr(), None); idle.append(time.perf_counter() - t0); ctx.synchronize()
for _ in range(5):
    ts = []
    for _ in range(12):
        t0 = time.perf_counter(); ver.run_device(d_p.data_ptr(), d_v.data_ptr(), d_out.data_ptr(), None); ts.append(time.perf_counter() - t0)
    ctx.synchronize(); queued.append(ts)
print(json.dumps({"idle_stream_us_per_call": round(1e6 * statistics.median(idle), 1), "queued_us_per_call_by_position": [round(1e6 * statistics.median(q[i] for q in queued), 1) for i in range(12)],
                  "note": "one call = 512-byte parameter upload + one CUDA-graph launch (13 kernel nodes, 2 branches) for 8 batches of 1024 proofs"}))
