"""Accept vs reject path of one blocking 1024-proof verification (host buffers): a batch with one corrupted proof goes through the
two-level fallback (32 chunk MSMs, then the 32 proofs of the failing chunk).  Usage: python benchmarks/reject_path.py [bad proofs]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bulletproofs_b200 as bp
import bench

nbad_list = [int(x) for x in sys.argv[1:]] or [0, 1, 3, 8, 64]
BATCH, N, M = 1024, 64, 1
orc, og, proofs, Vs = bench.make_workload(BATCH, 0, M)
plen = len(proofs) // BATCH
ctx = bp.Context(0); gens = bp.Gens(ctx, N, M); t = bp.Transcript(bench.LABEL); ot = orc.transcript(bench.LABEL)
ver = bp.BatchVerifier(ctx, gens, t, N, M, BATCH, 1)
out = {}
import random
rnd = random.Random(5)
for nbad in nbad_list:
    pb = bytearray(proofs)
    bad = sorted(rnd.sample(range(BATCH), nbad))
    for i in bad:
        pb[i * plen + 140] ^= 1
    pb = bytes(pb)
    got = bp.verify_batch(ctx, gens, t, pb, Vs, N, M, BATCH)
    assert [i for i, v in enumerate(got) if v] == bad
    best = 1e9
    for _ in range(8):
        t0 = time.perf_counter(); bp.verify_batch(ctx, gens, t, pb, Vs, N, M, BATCH); best = min(best, time.perf_counter() - t0)
    out[f"{nbad}_bad"] = round(best * 1e3, 3)
print(json.dumps({"blocking_call_ms_1024_proofs": out}))
