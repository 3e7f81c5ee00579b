"""Pure-Python big-integer restatement of ristretto255 (SURVEY.md §7 step 1 / §8c recipe): decode, encode, Elligator, addition, scalar
multiplication with Python ints only.  Slow by design; used by tests/test_pyref.py as an INDEPENDENT check of the C oracle's group layer
(the C oracle in turn checks the CUDA path).  Test infrastructure only."""
P = 2**255 - 19
L = 2**252 + 27742317777372353535851937790883648493
D = (-121665 * pow(121666, P - 2, P)) % P
SQRT_M1 = pow(2, (P - 1) // 4, P)
INVSQRT_A_MINUS_D = 54469307008909316920995813868745141605393597292927456921205312896311721017578
SQRT_AD_MINUS_ONE = 25063068953384623474111414158702152701244531502492656460079210482610430750235
ONE_MINUS_D_SQ = (1 - D * D) % P
D_MINUS_ONE_SQ = (D - 1) * (D - 1) % P


def is_neg(x):
    return (x % P) & 1


def fabs(x):
    x %= P
    return P - x if x & 1 else x


def sqrt_ratio_i(u, v):
    u %= P; v %= P
    v3 = v * v % P * v % P; v7 = v3 * v3 % P * v % P
    r = u * v3 % P * pow(u * v7 % P, (P - 5) // 8, P) % P
    chk = v * r % P * r % P
    ok = chk == u; f = chk == (-u) % P; fi = chk == (-u) * SQRT_M1 % P
    if f or fi:
        r = r * SQRT_M1 % P
    return (ok or f), fabs(r)


def decode(b):
    s = int.from_bytes(b, "little")
    if s >= P or s & 1:
        return None
    ss = s * s % P; u1 = (1 - ss) % P; u2 = (1 + ss) % P; u2s = u2 * u2 % P
    v = (-D * u1 % P * u1 - u2s) % P
    ok, I = sqrt_ratio_i(1, v * u2s % P)
    dx = I * u2 % P; dy = I * dx % P * v % P
    x = fabs(2 * s * dx % P); y = u1 * dy % P; t = x * y % P
    if not ok or is_neg(t) or y == 0:
        return None
    return (x, y, 1, t)


def encode(pt):
    X, Y, Z, T = pt
    u1 = (Z + Y) * (Z - Y) % P; u2 = X * Y % P
    _, I = sqrt_ratio_i(1, u1 * u2 % P * u2 % P)
    d1 = I * u1 % P; d2 = I * u2 % P; zi = d1 * d2 % P * T % P
    if is_neg(T * zi % P):
        x, y, di = Y * SQRT_M1 % P, X * SQRT_M1 % P, d1 * INVSQRT_A_MINUS_D % P
    else:
        x, y, di = X, Y, d2
    if is_neg(x * zi % P):
        y = (-y) % P
    return fabs(di * (Z - y) % P).to_bytes(32, "little")


def add(p, q):
    X1, Y1, Z1, T1 = p; X2, Y2, Z2, T2 = q
    A = (Y1 - X1) * (Y2 - X2) % P; B = (Y1 + X1) * (Y2 + X2) % P; C = 2 * D * T1 % P * T2 % P; Dd = 2 * Z1 * Z2 % P
    E, F, G, H = (B - A) % P, (Dd - C) % P, (Dd + C) % P, (B + A) % P
    return (E * F % P, G * H % P, F * G % P, E * H % P)


IDENT = (0, 1, 1, 0)


def mul(k, p):
    r = IDENT
    for bit in bin(k)[2:] if k else "":
        r = add(r, r)
        if bit == "1":
            r = add(r, p)
    return r


def msm(scalars, points):
    r = IDENT
    for k, p in zip(scalars, points):
        r = add(r, mul(k, p))
    return r


def elligator(r0):
    r = SQRT_M1 * r0 % P * r0 % P
    u = (r + 1) * ONE_MINUS_D_SQ % P; v = (-1 - r * D) * (r + D) % P
    ok, s = sqrt_ratio_i(u, v)
    sp = (-fabs(s * r0 % P)) % P
    c = -1
    if not ok:
        s, c = sp, r
    N = (c * (r - 1) % P * D_MINUS_ONE_SQ - v) % P
    w0, w1, w2, w3 = 2 * s * v % P, N * SQRT_AD_MINUS_ONE % P, (1 - s * s) % P, (1 + s * s) % P
    return (w0 * w3 % P, w2 * w1 % P, w1 * w3 % P, w0 * w2 % P)


def from_uniform(b64):
    r1 = int.from_bytes(b64[:32], "little") & (2**255 - 1); r2 = int.from_bytes(b64[32:], "little") & (2**255 - 1)
    return add(elligator(r1 % P), elligator(r2 % P))
