"""Host-side time of the C++ mirror alone (TEST-ONLY, CPU): runs the R1CS shuffle prover / verifier of bulletproofs_b200/host against the
mock engine and subtracts the time spent inside the engine's entry points (mock_engine_seconds), leaving what the mirror itself costs
per call -- constraint evaluation, polynomial algebra, transcripts, packing.  Usage: python tests/mock_engine/profile_mirror.py [k=4097] [reps=2]
With BP_MOCK_SKIP_POINT_MATH=1 in the environment the mock returns its first input point instead of computing (the proofs then do not verify):
the run finishes in seconds at 2^16 multipliers and a profiler sees the mirror, not the mock's curve arithmetic."""
import ctypes, json, os, random, sys, time
HERE = os.path.dirname(os.path.abspath(__file__)); ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bulletproofs_b200 as bp
bp.LIB_PATH = os.path.join(HERE, "libmockbpmsm.so"); bp.HOST_LIB_PATH = os.path.join(HERE, "libbulletproofs_host_mock.so")
from oracle_binding import L_ORDER

k = int(sys.argv[1]) if len(sys.argv) > 1 else 4097
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
mock = ctypes.CDLL(bp.LIB_PATH); mock.mock_engine_seconds.restype = ctypes.c_double
cap = 1 << (2 * (k - 1) - 1).bit_length()
ctx = bp.Context(0); gens = bp.Gens(ctx, cap, 1)
rnd = random.Random(5)
inp = [rnd.randrange(1 << 64) for _ in range(k)]; out = inp[:]; rnd.shuffle(out)
bl = b"".join(rnd.randrange(L_ORDER).to_bytes(32, "little") for _ in range(2 * k))

def tr():
    t = bp.Transcript(b"ShuffleBenchmark"); t.append_message(b"dom-sep", b"ShuffleProof"); t.append_u64(b"k", k); return t

res = {"k": k, "multipliers": 2 * (k - 1), "runs": []}
for rep in range(reps):
    e0 = mock.mock_engine_seconds(); t0 = time.perf_counter()
    rc, proof, V = bp.r1cs_prove(ctx, gens, tr(), bp.GADGET_SHUFFLE, inp + out, bl)
    wall_p = time.perf_counter() - t0; eng_p = mock.mock_engine_seconds() - e0
    e0 = mock.mock_engine_seconds(); t0 = time.perf_counter()
    ok = bp.r1cs_verify(ctx, gens, tr(), bp.GADGET_SHUFFLE, V, proof)
    wall_v = time.perf_counter() - t0; eng_v = mock.mock_engine_seconds() - e0
    assert rc == 0 and (ok == 0 or os.environ.get("BP_MOCK_SKIP_POINT_MATH"))
    res["runs"].append({"prove_host_s": round(wall_p - eng_p, 4), "prove_engine_s": round(eng_p, 3), "verify_host_s": round(wall_v - eng_v, 4), "verify_engine_s": round(eng_v, 3)})
print(json.dumps(res))
