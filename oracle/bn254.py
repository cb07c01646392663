"""BN254 ("bn128") field and G1 arithmetic in Python big integers.

Oracle restatement of ffjavascript 0.2.48 / wasmcurves 0.1.0 semantics
(un-vendored: /root/reference/yarn.lock:3905,8173) as recalled in SURVEY.md
Appendix A.1.  Test infrastructure only.

Conventions (the same as the iden3 tool-chain):
  * Fr / Fq elements are plain ints in [0, modulus).
  * "LEM" = 32-byte little-endian Montgomery form (value * 2^256 mod m), the
    in-memory / in-zkey representation.
  * "LE"  = 32-byte little-endian canonical form (the .wtns representation).
  * "BE"  = 32-byte big-endian canonical (transcript / toRprUncompressed).
  * G1 points are affine tuples (x, y) of ints, or None for infinity.
"""

R_MOD = 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001  # Fr
P_MOD = 0x30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47  # Fq
MONT_R = 1 << 256
G1_GEN = (1, 2)
CURVE_B = 3
FR_S = 28  # 2-adicity of r - 1
FR_NQR = 5  # smallest quadratic non-residue, ffjavascript picks it from 2 upward

assert (R_MOD - 1) % (1 << FR_S) == 0 and ((R_MOD - 1) >> FR_S) & 1


def fr_inv(a):
    return pow(a, R_MOD - 2, R_MOD)


def fq_inv(a):
    return pow(a, P_MOD - 2, P_MOD)


# --- roots of unity: w[28] = 5^((r-1)/2^28), w[i] = w[i+1]^2 ---
_W = [0] * (FR_S + 1)
_W[FR_S] = pow(FR_NQR, (R_MOD - 1) >> FR_S, R_MOD)
for _i in range(FR_S - 1, -1, -1):
    _W[_i] = _W[_i + 1] * _W[_i + 1] % R_MOD
assert _W[0] == 1 and _W[1] == R_MOD - 1


def fr_root(log_n):
    """Primitive 2^log_n-th root of unity used by Fr.fft."""
    return _W[log_n]


# --- byte codecs ---
def to_lem(a, m=R_MOD):
    return ((a << 256) % m).to_bytes(32, "little")


def from_lem(b, m=R_MOD):
    return int.from_bytes(b, "little") * pow(MONT_R, -1, m) % m


def to_le(a):
    return a.to_bytes(32, "little")


def from_le(b):
    return int.from_bytes(b, "little")


def to_be(a):
    return a.to_bytes(32, "big")


def g1_to_lem(P):
    """zkey / ptau in-file form: x||y, Fq Montgomery LE; infinity = zeros."""
    if P is None:
        return bytes(64)
    return to_lem(P[0], P_MOD) + to_lem(P[1], P_MOD)


def g1_from_lem(b):
    if b == bytes(64):
        return None
    return (from_lem(b[:32], P_MOD), from_lem(b[32:64], P_MOD))


def g1_to_be(P):
    """toRprUncompressed: x||y big-endian canonical; infinity = zeros."""
    if P is None:
        return bytes(64)
    return to_be(P[0]) + to_be(P[1])


# --- G1, Jacobian internally ---
def g1_is_on_curve(P):
    if P is None:
        return True
    x, y = P
    return (y * y - x * x * x - CURVE_B) % P_MOD == 0


def _jac_double(X, Y, Z):
    if Z == 0:
        return (1, 1, 0)
    p = P_MOD
    A = X * X % p
    B = Y * Y % p
    C = B * B % p
    D = 2 * ((X + B) * (X + B) - A - C) % p
    E = 3 * A % p
    F = E * E % p
    X3 = (F - 2 * D) % p
    Y3 = (E * (D - X3) - 8 * C) % p
    Z3 = 2 * Y * Z % p
    return (X3, Y3, Z3)


def _jac_add(P, Q):
    X1, Y1, Z1 = P
    X2, Y2, Z2 = Q
    if Z1 == 0:
        return Q
    if Z2 == 0:
        return P
    p = P_MOD
    Z1Z1 = Z1 * Z1 % p
    Z2Z2 = Z2 * Z2 % p
    U1 = X1 * Z2Z2 % p
    U2 = X2 * Z1Z1 % p
    S1 = Y1 * Z2 * Z2Z2 % p
    S2 = Y2 * Z1 * Z1Z1 % p
    if U1 == U2:
        if S1 == S2:
            return _jac_double(X1, Y1, Z1)
        return (1, 1, 0)
    H = (U2 - U1) % p
    I = 4 * H * H % p
    J = H * I % p
    r = 2 * (S2 - S1) % p
    V = U1 * I % p
    X3 = (r * r - J - 2 * V) % p
    Y3 = (r * (V - X3) - 2 * S1 * J) % p
    Z3 = ((Z1 + Z2) * (Z1 + Z2) - Z1Z1 - Z2Z2) * H % p
    return (X3, Y3, Z3)


def _to_jac(P):
    return (1, 1, 0) if P is None else (P[0], P[1], 1)


def _to_aff(J):
    X, Y, Z = J
    if Z == 0:
        return None
    zi = fq_inv(Z)
    zi2 = zi * zi % P_MOD
    return (X * zi2 % P_MOD, Y * zi2 * zi % P_MOD)


def g1_add(P, Q):
    return _to_aff(_jac_add(_to_jac(P), _to_jac(Q)))


def g1_neg(P):
    return None if P is None else (P[0], (-P[1]) % P_MOD)


def g1_sub(P, Q):
    return g1_add(P, g1_neg(Q))


def g1_mul(P, k):
    k %= R_MOD
    acc = (1, 1, 0)
    base = _to_jac(P)
    while k:
        if k & 1:
            acc = _jac_add(acc, base)
        base = _jac_double(*base)
        k >>= 1
    return _to_aff(acc)


def g1_msm_naive(points, scalars):
    """sum_i scalars[i] * points[i]; the definition G1.multiExpAffine computes."""
    acc = (1, 1, 0)
    for P, k in zip(points, scalars):
        if P is None or k % R_MOD == 0:
            continue
        acc = _jac_add(acc, _to_jac(g1_mul(P, k)))
    return _to_aff(acc)


def g1_msm(points, scalars, c=None):
    """Pippenger bucket method (same result as g1_msm_naive, faster for big n)."""
    n = len(points)
    if n == 0:
        return None
    if c is None:
        c = max(2, min(16, n.bit_length() - 2))
    nwin = (254 + c - 1) // c
    total = (1, 1, 0)
    for w in range(nwin - 1, -1, -1):
        for _ in range(c):
            total = _jac_double(*total)
        buckets = [None] * (1 << c)
        sh = w * c
        mask = (1 << c) - 1
        for P, k in zip(points, scalars):
            if P is None:
                continue
            d = (k >> sh) & mask
            if d:
                b = buckets[d]
                buckets[d] = _to_jac(P) if b is None else _jac_add(b, _to_jac(P))
        run = (1, 1, 0)
        acc = (1, 1, 0)
        for d in range((1 << c) - 1, 0, -1):
            if buckets[d] is not None:
                run = _jac_add(run, buckets[d])
            acc = _jac_add(acc, run)
        total = _jac_add(total, acc)
    return _to_aff(total)


def srs_g1(tau, count):
    """[tau^i]G1 for i < count (insecure, known-trapdoor SRS; Makefile:64-67 role)."""
    out = []
    t = 1
    for _ in range(count):
        out.append(g1_mul(G1_GEN, t))
        t = t * tau % R_MOD
    return out
