"""Small end-to-end pass for compute-sanitizer (memcheck / racecheck / initcheck): batch verification with valid, damaged and
malformed proofs (RLC path + per-proof fallback), an MSM with heavy buckets, the IPP prover rounds."""
import os, random, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bulletproofs_b200 as bp
from oracle_binding import Oracle, L_ORDER

orc = Oracle(); ctx = bp.Context(0); rnd = random.Random(2)
label = b"sanitize"
for n, m, count in ((64, 1, 40), (16, 4, 9)):
    og = orc.gens(n, m); gens = bp.Gens(ctx, n, m)
    vals = [rnd.randrange(1 << n) for _ in range(count * m)]; bl = b"".join(rnd.randrange(L_ORDER).to_bytes(32, "little") for _ in range(count * m))
    seeds = b"".join(i.to_bytes(8, "little") + bytes(24) for i in range(count))
    proofs, Vs = orc.prove_many(og, orc.transcript(label), vals, bl, n, m, seeds, nthreads=4)
    plen = len(proofs) // count; t = bp.Transcript(label)
    assert bp.verify_batch(ctx, gens, t, proofs, Vs, n, m, count) == [0] * count
    bad = bytearray(proofs); bad[3 * plen + 130] ^= 1; bad[5 * plen:5 * plen + 32] = bytes(32); bad[7 * plen + 128:7 * plen + 160] = b"\xff" * 32
    got = bp.verify_batch(ctx, gens, t, bytes(bad), Vs, n, m, count)
    want = orc.verify_many(og, orc.transcript(label), bytes(bad), plen, Vs, n, m, count, nthreads=4)
    assert got == want, (got, want)
    rc, proof, V = bp.prove_multiple(ctx, gens, bp.Transcript(label), vals[:m], bl[:32 * m], n, bytes(32))
    assert rc == 0 and bp.verify_multiple(ctx, gens, bp.Transcript(label), proof, V, n) == 0
    gens.close()
base = [orc.from_uniform(rnd.randbytes(64)) for _ in range(32)]
nn = 4096
sc = b"".join(rnd.randrange(L_ORDER).to_bytes(32, "little") for _ in range(nn)); pp = b"".join(rnd.choice(base) for _ in range(nn))
assert ctx.msm(sc, pp) == orc.msm(sc, pp)
print("sanitize_smoke ok")
