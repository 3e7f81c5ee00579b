"""Multi-stream timeline of the config-2 verification step (diagnostic, no nsys in the image): S contexts/streams as in bench.py,
CUDA events around every launch, all timestamps relative to one reference event.  Prints, for the steady state, each kernel's
mean duration under contention next to its solo duration, and the time-weighted number of concurrently running kernels.
Usage: python benchmarks/timeline.py [streams] [rounds] [batches per group]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import numpy as np, torch
import bulletproofs_b200 as bp
import bench

S = int(sys.argv[1]) if len(sys.argv) > 1 else 24
ROUNDS = int(sys.argv[2]) if len(sys.argv) > 2 else 6
G = int(sys.argv[3]) if len(sys.argv) > 3 else 8
BATCH, N, M = 1024, 64, 1
orc, og, proofs, Vs = bench.make_workload(BATCH, 0, M)
plen = len(proofs) // BATCH
streams = [torch.cuda.Stream() for _ in range(S)]
ctxs = [bp.Context(0, stream=s.cuda_stream) for s in streams]
gens = [bp.Gens(c, N, M) for c in ctxs]
t = bp.Transcript(bench.LABEL)
ver = [bp.BatchVerifier(ctxs[i], gens[i], t, N, M, BATCH, G) for i in range(S)]
P = 24
pr = np.frombuffer(proofs, dtype=np.uint8).reshape(BATCH, plen); vs = np.frombuffer(Vs, dtype=np.uint8).reshape(BATCH, 32)
d_proofs = torch.stack([torch.cat([torch.from_numpy(np.roll(pr, (i * G + j) * 5, axis=0).copy()) for j in range(G)]) for i in range(P)]).cuda()
d_vs = torch.stack([torch.cat([torch.from_numpy(np.roll(vs, (i * G + j) * 5, axis=0).copy()) for j in range(G)]) for i in range(P)]).cuda()
d_verdicts = torch.zeros((S, G * BATCH), dtype=torch.int32, device="cuda")

def step(i):
    k, j = i % S, i % P
    ver[k].run_device(d_proofs[j].data_ptr(), d_vs[j].data_ptr(), d_verdicts[k].data_ptr(), None)

for i in range(3 * S): step(i)
torch.cuda.synchronize()
# solo durations
ctxs[0].prof_enable(True)
for i in range(5): ver[0].run_device(d_proofs[i].data_ptr(), d_vs[i].data_ptr(), d_verdicts[0].data_ptr(), None)
solo = {k: v[0] / v[1] for k, v in ctxs[0].prof_report().items()}       # ms per launch
ctxs[0].prof_enable(False)
for c in ctxs: c.prof_enable(True)
for i in range(ROUNDS * S): step(i)
torch.cuda.synchronize()
recs = []
for si, c in enumerate(ctxs):
    recs += [(si, name, a, b) for name, a, b in c.prof_timeline(ctxs[0])]
for c in ctxs: c.prof_enable(False)
t_end = max(r[3] for r in recs); t_beg = min(r[2] for r in recs)
lo, hi = t_beg + 0.25 * (t_end - t_beg), t_beg + 0.75 * (t_end - t_beg)         # steady-state window
mid = [r for r in recs if r[2] >= lo and r[3] <= hi]
per = {}
for _, name, a, b in mid:
    d = per.setdefault(name, [0.0, 0]); d[0] += b - a; d[1] += 1
batches = G * sum(1 for r in mid if r[1] == "k_rp_transcript")
out = {"streams": S, "batches_per_group": G, "window_ms": hi - lo, "batches_in_window": batches, "us_per_batch": 1e3 * (hi - lo) / max(1, batches), "kernels": {}}
for name, (tot, cnt) in sorted(per.items(), key=lambda kv: -kv[1][0]):
    out["kernels"][name] = {"mean_us_contended": round(1e3 * tot / cnt, 1), "solo_us": round(1e3 * solo.get(name, 0.0), 1),
                            "launches": cnt, "avg_concurrent": round(tot / (hi - lo), 2)}
out["avg_kernels_running"] = round(sum(v["avg_concurrent"] for v in out["kernels"].values()), 2)
print(json.dumps(out))
