"""BASELINE config 5 (secondary result, not the bench line): k-shuffle with 2^16 multipliers, prove + verify,
GPU-backed path vs the CPU oracle.  Writes one JSON line."""
import json, os, random, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bulletproofs_b200 as bp
from oracle_binding import Oracle, L_ORDER

k = int(sys.argv[1]) if len(sys.argv) > 1 else 32769
cpu = "--cpu" in sys.argv
cap = 1 << (2 * (k - 1) - 1).bit_length()
orc = Oracle(); ctx = bp.Context(0)
t0 = time.perf_counter(); gens = bp.Gens(ctx, cap, 1); t_gens = time.perf_counter() - t0
rnd = random.Random(5)
inp = [rnd.randrange(1 << 64) for _ in range(k)]; out = inp[:]; rnd.shuffle(out)
bl = b"".join(rnd.randrange(L_ORDER).to_bytes(32, "little") for _ in range(2 * k))

def tr():
    t = bp.Transcript(b"ShuffleBenchmark"); t.append_message(b"dom-sep", b"ShuffleProof"); t.append_u64(b"k", k); return t

res = {"config": f"R1CS shuffle, k={k}, {2 * (k - 1)} multipliers, gens capacity {cap}", "gens_table_s": round(t_gens, 3)}
for rep in range(2):
    t0 = time.perf_counter(); rc, proof, V = bp.r1cs_prove(ctx, gens, tr(), bp.GADGET_SHUFFLE, inp + out, bl); t_prove = time.perf_counter() - t0
    t0 = time.perf_counter(); ok = bp.r1cs_verify(ctx, gens, tr(), bp.GADGET_SHUFFLE, V, proof); t_verify = time.perf_counter() - t0
    assert rc == 0 and ok == 0
res.update({"gpu_prove_s": round(t_prove, 4), "gpu_verify_s": round(t_verify, 4), "proof_bytes": len(proof), "verify_msm_terms": 13 + 2 * k + 2 * cap + 2 * (cap.bit_length() - 1)})
if cpu:
    og = orc.gens(cap, 1)
    ot = orc.transcript(b"ShuffleBenchmark"); ot = orc.transcript_append(ot, b"dom-sep", b"ShuffleProof"); ot = orc.transcript_append(ot, b"k", k.to_bytes(8, "little"))
    t0 = time.perf_counter(); assert orc.r1cs_verify(og, ot, 0, V, proof) == 0; res["cpu_oracle_verify_s"] = round(time.perf_counter() - t0, 3)
    t0 = time.perf_counter(); orc_rc, want, _ = orc.r1cs_prove(og, ot, 0, inp + out, bl); res["cpu_oracle_prove_s"] = round(time.perf_counter() - t0, 3)
    res["proof_bytes_equal_oracle"] = (want == proof)
print(json.dumps(res))
