"""GPU tier: the FP64-pipe form of the decompression ladder (csrc/fd.cuh) gives the same bits as the integer form and as the oracle."""
import random

import pytest

from oracle_binding import L_ORDER as l, P_FIELD as p
from test_gpu_parity import _workload, le

pytestmark = pytest.mark.gpu


def test_fp64_field_arithmetic_against_bigints(gpu_ctx):
    rnd = random.Random(11)
    ones43 = sum(((1 << 43) - 1) << s for s in (0, 85, 170)); ones42 = sum(((1 << 42) - 1) << s for s in (43, 128, 213))
    edge = [0, 1, 2, 19, p - 1, p, p + 1, 2**255 - 1, 2**255, 2**255 + 18, 2**256 - 1, 2**256 - 38, 2**43 - 1, 2**43, 2**85 - 1, ones43, ones42, ones43 | ones42]
    vals = edge + [rnd.getrandbits(256) for _ in range(2000)] + [rnd.getrandbits(256) | (2**256 - 2**200) for _ in range(200)] + [rnd.getrandbits(64) for _ in range(100)]
    A = [rnd.choice(vals) for _ in range(20000)] + [a for a in edge for _ in edge]
    B = [rnd.choice(vals) for _ in range(20000)] + [b for _ in edge for b in edge]
    a, b = b"".join(le(x) for x in A), b"".join(le(x) for x in B)
    fns = {8: lambda x, y: x * y % p, 9: lambda x, y: x * x % p}
    for op, f in fns.items():
        out = gpu_ctx.debug_fe_op(op, a, b)
        for i, (x, y) in enumerate(zip(A, B)):
            assert int.from_bytes(out[32 * i:32 * i + 32], "little") == f(x, y), (op, hex(x), hex(y))
    sel = list(range(600))
    a2 = b"".join(le(A[i]) for i in sel) + b"".join(le(x) for x in edge); b2 = b"".join(le(B[i]) for i in sel) + b"".join(le(x) for x in reversed(edge))
    A2 = [A[i] for i in sel] + edge; B2 = [B[i] for i in sel] + list(reversed(edge))
    inv, pw, ch = gpu_ctx.debug_fe_op(10, a2, b2), gpu_ctx.debug_fe_op(11, a2, b2), gpu_ctx.debug_fe_op(12, a2, b2)
    for i, (x, y) in enumerate(zip(A2, B2)):
        assert int.from_bytes(inv[32 * i:32 * i + 32], "little") == pow(x % p, p - 2, p), hex(x)
        assert int.from_bytes(pw[32 * i:32 * i + 32], "little") == pow(x % p, (p - 5) // 8, p), hex(x)
        assert int.from_bytes(ch[32 * i:32 * i + 32], "little") == pow(pow(x * y, 2, p) * x, 2**20, p) * y % p, (hex(x), hex(y))


@pytest.mark.parametrize("share", [8, 3])
def test_fp64_decompression_matches_oracle_and_integer_path(orc, share):
    import bulletproofs_b200 as bp
    ctx = bp.Context(0)
    rnd = random.Random(12)
    pts = [orc.from_uniform(rnd.randbytes(64)) for _ in range(400)]
    cand = list(pts)
    for i in range(4000):
        s = bytearray(rnd.randbytes(32))
        if i % 4 == 0: s[31] |= 0x80
        if i % 5 == 0:
            s = bytearray(rnd.choice(pts)); s[rnd.randrange(32)] ^= 1 << rnd.randrange(8)
        cand.append(bytes(s))
    cand += [bytes(32), le(p), le(p + 2), le(1), le(2**255 - 20), b"\xff" * 32]
    blob = b"".join(cand)
    want_ok = [orc.point_is_valid(c) for c in cand]
    vals0, ok0 = ctx.decompress(blob)
    ctx.set_fp64_share(share)
    assert ctx.decompress_check(blob) == want_ok
    vals1, ok1 = ctx.decompress(blob)
    assert ok1 == ok0 == want_ok and vals1 == vals0
    # an MSM whose points are decompressed inside the call, and a resident point set built on the FP64 path
    sc = b"".join(le(rnd.randrange(l)) for _ in range(400))
    assert ctx.msm(sc, b"".join(pts)) == orc.msm(sc, b"".join(pts))
    ps = bp.PointSet(ctx, b"".join(pts))
    st, res = ps.msm(sc, 1, 400)
    assert st == [0] and res[0] == orc.msm(sc, b"".join(pts))[1]
    ps.close(); ctx.close()


@pytest.mark.parametrize("share", [8, 5])
def test_fp64_share_batch_verify_matches_oracle(orc, share):
    """per-proof verdicts == oracle with the ladder on the FP64 pipe: direct launches, a reserved (graph) geometry, damaged proofs"""
    import bulletproofs_b200 as bp
    ctx = bp.Context(0); ctx.set_fp64_share(share)
    label = b"AggregateRangeProofBenchmark"
    n, m, count, nb = 64, 1, 40, 3
    og = orc.gens(64, 1); gens = bp.Gens(ctx, 64, 1)
    proofs, Vs = _workload(orc, og, label, n, m, count * nb, seed=600 + share)
    plen = len(proofs) // (count * nb)
    t = bp.Transcript(label); ot = orc.transcript(label)
    got, ok = bp.verify_group(ctx, gens, t, proofs, Vs, n, m, count, nb)
    assert got == [0] * (count * nb) and ok == [1] * nb
    pb, vb = bytearray(proofs), bytearray(Vs)
    pb[(0 * count + 7) * plen + 300] ^= 1
    pb[(1 * count + 3) * plen + 224:(1 * count + 3) * plen + 256] = le(p + 1)        # L_0 undecodable
    pb[(1 * count + 9) * plen + 32:(1 * count + 9) * plen + 64] = le(2)                # S = 2: a field element that is not a point
    vb[(2 * count + 39) * 32 + 5] ^= 8
    want = orc.verify_many(og, ot, bytes(pb), plen, bytes(vb), n, m, count * nb)
    for reserve in (False, True):
        if reserve:
            v = bp.BatchVerifier(ctx, gens, t, n, m, count, nb)
        got, ok = bp.verify_group(ctx, gens, t, bytes(pb), bytes(vb), n, m, count, nb, seed=bytes([share]) * 32)
        assert got == want, (reserve, [(i, g, w) for i, (g, w) in enumerate(zip(got, want)) if g != w])
        assert ok == [0, 0, 0] and sum(1 for w in want if w) == 4
    # changing the share after the reservation: the stale graph is not replayed, verdicts stay right
    ctx.set_fp64_share(0)
    got, ok = bp.verify_group(ctx, gens, t, proofs, Vs, n, m, count, nb)
    assert got == [0] * (count * nb) and ok == [1] * nb
    gens.close(); ctx.close()
