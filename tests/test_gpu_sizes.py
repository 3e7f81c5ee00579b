"""GPU tier: BASELINE configs 3 and 4 at full size (parity-test cases, not bench lines).

config 3: 256 aggregated (64,16) proofs per batch;  config 4: Ristretto MSM size sweep n = 2^10 .. 2^20.
Exact oracle comparison where the oracle finishes in seconds, size-independent exact properties beyond that:
with points tiled from a small base set, sum_i s_i P_(i mod b) = sum_j (sum_{i = j mod b} s_i) P_j, so the
2^20-term result must equal a b-term oracle MSM over the folded scalars."""
import os
import random

import pytest

from oracle_binding import L_ORDER as l

pytestmark = pytest.mark.gpu


def le(x, n=32):
    return x.to_bytes(n, "little")


@pytest.mark.parametrize("lg", [10, 12, 14, 16])
def test_msm_sweep_exact(gpu_ctx, orc, lg):
    n = 1 << lg
    rnd = random.Random(lg)
    base = [orc.from_uniform(rnd.randbytes(64)) for _ in range(256)]
    sc = b"".join(le(rnd.randrange(l)) for _ in range(n))
    pp = b"".join(base[i % 256] for i in range(n))
    assert gpu_ctx.msm(sc, pp) == orc.msm(sc, pp)


@pytest.mark.parametrize("lg", [18, 20])
def test_msm_sweep_folded_property(gpu_ctx, orc, lg):
    n, b = 1 << lg, 1024
    rnd = random.Random(lg)
    base = [orc.from_uniform(rnd.randbytes(64)) for _ in range(b)]
    scalars = [rnd.randrange(l) for _ in range(n)]
    folded = [0] * b
    for i, s in enumerate(scalars):
        folded[i % b] += s
    sc = b"".join(le(s) for s in scalars)
    pp = b"".join(base) * (n // b)
    want = orc.msm(b"".join(le(f % l) for f in folded), b"".join(base))
    assert gpu_ctx.msm(sc, pp) == want


def test_msm_batch_of_eight_per_size(gpu_ctx, orc):
    """config 4's shape: a batch of >= 8 independent MSMs of one size in one call."""
    rnd = random.Random(44)
    base = [orc.from_uniform(rnd.randbytes(64)) for _ in range(128)]
    n = 1 << 11
    sc = b"".join(le(rnd.randrange(l)) for _ in range(8 * n)); pp = b"".join(rnd.choice(base) for _ in range(8 * n))
    status, outs = gpu_ctx.msm_batch(sc, pp, [i * n for i in range(9)])
    for j in range(8):
        assert (status[j], outs[j]) == orc.msm(sc[32 * j * n:32 * (j + 1) * n], pp[32 * j * n:32 * (j + 1) * n])


def test_config3_full_size_batch(gpu_ctx, orc):
    """256 x (64,16): 16 oracle-made aggregated proofs tiled 16 times verify; single damaged proofs are located exactly;
    the first 16 verdicts equal the oracle's per-proof verdicts."""
    import bulletproofs_b200 as bp
    label = b"AggregateRangeProofBenchmark"
    n, m, base_count, count = 64, 16, 16, 256
    og = orc.gens(64, 16); gens = bp.Gens(gpu_ctx, 64, 16)
    rnd = random.Random(316)
    values = [rnd.randrange(1 << n) for _ in range(base_count * m)]
    blind = b"".join(le(rnd.randrange(l)) for _ in range(base_count * m))
    seeds = b"".join(le(i, 8) + bytes(24) for i in range(base_count))
    proofs, Vs = orc.prove_many(og, orc.transcript(label), values, blind, n, m, seeds, nthreads=os.cpu_count() or 4)
    plen = len(proofs) // base_count
    assert plen == 928
    big_p, big_v = proofs * (count // base_count), Vs * (count // base_count)
    t = bp.Transcript(label)
    assert bp.verify_batch(gpu_ctx, gens, t, big_p, big_v, n, m, count) == [0] * count
    pb = bytearray(big_p); vb = bytearray(big_v)
    pb[3 * plen + 500] ^= 1; vb[(200 * m + 7) * 32 + 1] ^= 8
    got = bp.verify_batch(gpu_ctx, gens, t, bytes(pb), bytes(vb), n, m, count)
    assert [i for i, v in enumerate(got) if v] == [3, 200]
    assert got[:16] == orc.verify_many(og, orc.transcript(label), bytes(pb[:16 * plen]), plen, bytes(vb[:16 * 32 * m]), n, m, 16)
    gens.close()


def test_resident_point_set_large(gpu_ctx, orc):
    """bp_points at sweep size: 2 MSMs of 2^18 terms over one resident set (points tiled from 512 bases, so the result equals a 512-term
    oracle MSM over the folded scalars); the decompress-every-call path gives the same bytes."""
    import bulletproofs_b200 as bp
    n, b, K = 1 << 18, 512, 2
    rnd = random.Random(1818)
    base = [orc.from_uniform(rnd.randbytes(64)) for _ in range(b)]
    pts = b"".join(base) * (n // b)
    ps = bp.PointSet(gpu_ctx, pts)
    scalars = [[rnd.randrange(l) for _ in range(n)] for _ in range(K)]
    sc = b"".join(le(s) for row in scalars for s in row)
    st, outs = ps.msm(sc, K, n)
    assert st == [0] * K
    for j in range(K):
        folded = [0] * b
        for i, s in enumerate(scalars[j]):
            folded[i % b] += s
        assert (0, outs[j]) == orc.msm(b"".join(le(f % l) for f in folded), b"".join(base)), j
    assert gpu_ctx.msm(sc[:32 * n], pts) == (0, outs[0])
    ps.close()
