"""Helper for ncu: one bp_msm_batch call of 32 MSMs x 17 538 terms (the config-2 MSM shape at full occupancy),
so that the wide kernels can be profiled with every SM filled."""
import os, sys, random
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bulletproofs_b200 as bp
from oracle_binding import Oracle, L_ORDER

orc = Oracle(); ctx = bp.Context(0); rnd = random.Random(1)
base = [orc.from_uniform(rnd.randbytes(64)) for _ in range(512)]
n, n_msm = 17538, 32
T = n * n_msm
sc = b"".join(rnd.randrange(L_ORDER).to_bytes(32, "little") for _ in range(T)); pp = b"".join(base[i % 512] for i in range(T))
offs = [i * n for i in range(n_msm + 1)]
for _ in range(2):
    st, outs = ctx.msm_batch(sc, pp, offs)
assert not any(st)
