import os, sys, random, time
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import bulletproofs_b200 as bp
from oracle_binding import Oracle, L_ORDER
orc = Oracle(); ctx = bp.Context(0); rnd = random.Random(7)
basep = [orc.from_uniform(rnd.randbytes(64)) for _ in range(1024)]
for lg, k in ((12, 8), (14, 8), (16, 8), (18, 8)):
    nn = 1 << lg
    sc = b"".join(rnd.randrange(L_ORDER).to_bytes(32, "little") for _ in range(nn)) * k
    pp = (b"".join(basep) * (nn // 1024)) * k
    offs = [i * nn for i in range(k + 1)]
    ctx.msm_batch(sc, pp, offs)
    ctx.prof_enable(True)
    t0 = time.perf_counter(); ctx.msm_batch(sc, pp, offs); dt = time.perf_counter() - t0
    rep = ctx.prof_report(); ctx.prof_enable(False)
    print(lg, round(dt * 1e3, 2), {k_: round(v[0], 3) for k_, v in rep.items()})
# config 3 breakdown
label = b"AggregateRangeProofBenchmark"; n, m, base, count = 64, 16, 16, 256
og = orc.gens(64, 16); gens = bp.Gens(ctx, 64, 16)
vals = [rnd.randrange(1 << n) for _ in range(base * m)]; bl = b"".join(rnd.randrange(L_ORDER).to_bytes(32, "little") for _ in range(base * m))
seeds = b"".join(i.to_bytes(8, "little") + bytes(24) for i in range(base))
proofs, Vs = orc.prove_many(og, orc.transcript(label), vals, bl, n, m, seeds, nthreads=16)
P, V = proofs * (count // base), Vs * (count // base); t = bp.Transcript(label)
bp.verify_batch(ctx, gens, t, P, V, n, m, count)
ctx.prof_enable(True)
t0 = time.perf_counter(); bp.verify_batch(ctx, gens, t, P, V, n, m, count); dt = time.perf_counter() - t0
rep = ctx.prof_report(); ctx.prof_enable(False)
print("cfg3", round(dt * 1e3, 2), {k_: (round(v[0], 3), v[1]) for k_, v in rep.items()})
