"""Experiment: per-kernel time of the MSM pipeline for 1 vs many identical-size MSMs in one launch sequence
(how much faster the wide kernels run when the grid is large enough to fill the GPU)."""
import os, sys, random, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bulletproofs_b200 as bp
from oracle_binding import Oracle, L_ORDER

orc = Oracle(); ctx = bp.Context(0)
rnd = random.Random(1)
base = [orc.from_uniform(rnd.randbytes(64)) for _ in range(512)]
n = 17538
for n_msm in (1, 8, 32):
    T = n * n_msm
    sc = b"".join(rnd.randrange(L_ORDER).to_bytes(32, "little") for _ in range(T)); pp = b"".join(base[i % 512] for i in range(T))
    offs = [i * n for i in range(n_msm + 1)]
    ctx.msm_batch(sc, pp, offs)
    ctx.prof_enable(True)
    for _ in range(3): ctx.msm_batch(sc, pp, offs)
    rep = ctx.prof_report(); ctx.prof_enable(False)
    print(n_msm, {k: round(v[0] / v[1] / n_msm * 1e3, 1) for k, v in rep.items()}, "us per MSM")
