"""CPU tier: the C oracle's group layer against the pure-Python big-integer restatement (tests/pyref.py), independent of both C code bases."""
import hashlib
import random

import pyref
from oracle_binding import L_ORDER as l


def test_constants():
    assert pyref.SQRT_M1 == 19681161376707505956807079304988542015446066515923890162744021073123829784752
    assert pyref.D == 37095705934669439343138083508754565189542113879843219016388785533085940283555
    assert pyref.INVSQRT_A_MINUS_D ** 2 * (-1 - pyref.D) % pyref.P == 1
    assert pyref.SQRT_AD_MINUS_ONE ** 2 % pyref.P == (-pyref.D - 1) % pyref.P


def test_points_and_msm_against_python_ints(orc):
    rnd = random.Random(8)
    B, Bb = orc.pedersen()
    assert pyref.encode(pyref.decode(B)) == B
    # B_blinding = from_uniform(SHA3-512(B)) (generators.rs:48-51); G[0][0] from the SHAKE256 chain (generators.rs:62-104)
    assert pyref.encode(pyref.from_uniform(hashlib.sha3_512(B).digest())) == Bb
    g0 = hashlib.shake_256(b"GeneratorsChain" + b"G" + (0).to_bytes(4, "little")).digest(128)
    assert pyref.encode(pyref.from_uniform(g0[:64])).hex() == "fc3b25801422672a6a8d3adb5d8457d4301fe92324b4fc56ae934c8713ddfe2d"
    assert pyref.encode(pyref.from_uniform(g0[64:])).hex() == "ae817fdef62f713dd169dc8a26406f68be0bd3cd53652614636b0801567c4264"
    pts = []
    for _ in range(12):
        u = rnd.randbytes(64)
        enc = orc.from_uniform(u)
        assert pyref.encode(pyref.from_uniform(u)) == enc
        p = pyref.decode(enc); assert p is not None and pyref.encode(p) == enc
        pts.append((enc, p))
    for _ in range(10):
        (ea, pa), (eb, pb) = rnd.choice(pts), rnd.choice(pts)
        assert pyref.encode(pyref.add(pa, pb)) == orc.point_add(ea, eb)
    # decode accept/reject agrees on malformed encodings
    for i in range(300):
        s = bytearray(rnd.randbytes(32))
        if i % 3 == 0:
            s = bytearray(rnd.choice(pts)[0]); s[rnd.randrange(32)] ^= 1 << rnd.randrange(8)
        if i % 5 == 0:
            s[31] |= 0x80
        assert (pyref.decode(bytes(s)) is not None) == bool(orc.point_is_valid(bytes(s)))
    # multiscalar multiplication: Straus-size MSMs of the oracle == naive Python double-and-add
    for n in (1, 2, 5, 9):
        ks = [rnd.randrange(l) for _ in range(n)]; sel = [rnd.choice(pts) for _ in range(n)]
        want = pyref.encode(pyref.msm(ks, [p for _, p in sel]))
        assert orc.msm(b"".join(k.to_bytes(32, "little") for k in ks), b"".join(e for e, _ in sel)) == (0, want)
    # the 8 golden commitments' construction: v B + b B~ (tests/range_proof.rs:108-113) for one value
    k, bl = 5, rnd.randrange(l)
    want = pyref.encode(pyref.add(pyref.mul(k, pyref.decode(B)), pyref.mul(bl, pyref.decode(Bb))))
    assert orc.msm(k.to_bytes(32, "little") + bl.to_bytes(32, "little"), B + Bb) == (0, want)
