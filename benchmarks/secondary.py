"""Secondary measurements for BASELINE configs 3 and 4 on one B200 (not the bench line): batched verify of 256 aggregated
(64,16) proofs, and the Ristretto MSM size sweep n = 2^10 .. 2^20 (8 MSMs per call, compressed inputs, host buffers)."""
import json, os, random, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bulletproofs_b200 as bp
from oracle_binding import Oracle, L_ORDER

orc = Oracle(); ctx = bp.Context(0); rnd = random.Random(7); res = {}
# ---- config 3: 256 x (64,16)
label = b"AggregateRangeProofBenchmark"; n, m, base, count = 64, 16, 16, 256
og = orc.gens(64, 16); gens = bp.Gens(ctx, 64, 16)
vals = [rnd.randrange(1 << n) for _ in range(base * m)]; bl = b"".join(rnd.randrange(L_ORDER).to_bytes(32, "little") for _ in range(base * m))
seeds = b"".join(i.to_bytes(8, "little") + bytes(24) for i in range(base))
proofs, Vs = orc.prove_many(og, orc.transcript(label), vals, bl, n, m, seeds, nthreads=16)
P, V = proofs * (count // base), Vs * (count // base); t = bp.Transcript(label)
assert bp.verify_batch(ctx, gens, t, P, V, n, m, count) == [0] * count
t0 = time.perf_counter(); reps = 20
for _ in range(reps): bp.verify_batch(ctx, gens, t, P, V, n, m, count)
dt = (time.perf_counter() - t0) / reps
t0 = time.perf_counter(); st = orc.verify_many(og, orc.transcript(label), proofs, len(proofs) // base, Vs, n, m, base, nthreads=16); dtc = time.perf_counter() - t0
res["config3_256x(64,16)"] = {"gpu_single_call_ms": round(dt * 1e3, 3), "gpu_proofs_per_s_single_stream": round(count / dt), "values_per_s": round(count * m / dt),
                              "cpu_oracle_proofs_per_s_16_threads": round(base / dtc, 1)}
# ---- config 4: MSM sweep
basep = [orc.from_uniform(rnd.randbytes(64)) for _ in range(1024)]
sweep = {}
for lg in range(10, 21, 2):
    nn = 1 << lg; k = 8 if lg <= 18 else 2
    sc = b"".join(rnd.randrange(L_ORDER).to_bytes(32, "little") for _ in range(nn)) * k
    pp = (b"".join(basep) * (nn // 1024)) * k
    offs = [i * nn for i in range(k + 1)]
    ctx.msm_batch(sc, pp, offs)
    t0 = time.perf_counter(); ctx.msm_batch(sc, pp, offs); dt = time.perf_counter() - t0
    entry = {"msms_per_call": k, "ms_per_call": round(dt * 1e3, 2), "terms_per_s": round(k * nn / dt)}
    if lg <= 16:
        t0 = time.perf_counter(); orc.msm(sc[:32 * nn], pp[:32 * nn]); entry["cpu_oracle_1thread_terms_per_s"] = round(nn / (time.perf_counter() - t0))
    sweep[f"2^{lg}"] = entry
res["config4_msm_sweep(host buffers, decompress included)"] = sweep
print(json.dumps(res))
