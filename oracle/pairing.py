"""TEST INFRASTRUCTURE -- BN254 optimal ate pairing in Python big integers.

Oracle for the device verifier (csrc/pairing.cuh, csrc/verify.cu).  Restates what snarkjs 0.4.12 `plonk.verify`
reaches through `curve.pairingEq` (ffjavascript 0.2.48 -> wasmcurves 0.1.0 `bn128_pairingEq*`; un-vendored,
/root/reference/yarn.lock:3905,8173): the optimal ate pairing on alt_bn128 with the D-type sextic twist,
Fq2 = Fq[u]/(u^2+1), Fq12 = Fq2[w]/(w^6 - xi), xi = 9 + u.  **Parity unpinned** against wasmcurves (no Node in
the image); pinned by the defining properties instead -- bilinearity, non-degeneracy, order r -- and by the
known-trapdoor G1 form of the PLONK check (oracle/plonk.py verify_with_trapdoor).

Written for transparency, not speed: Fq12 is a flat list of six Fq2 coefficients of 1, w, .., w^5 with schoolbook
products; G2 arithmetic is affine; the hard part of the final exponentiation is a plain power.
"""
from .bn254 import P_MOD as P, R_MOD as R

XI = (9, 1)
BN_X = 4965661367192848881                      # the BN parameter: p = 36x^4 + 36x^3 + 24x^2 + 6x + 1
ATE_LOOP = 6 * BN_X + 2                         # 29793968203157093288
assert 36 * BN_X**4 + 36 * BN_X**3 + 24 * BN_X**2 + 6 * BN_X + 1 == P
assert 36 * BN_X**4 + 36 * BN_X**3 + 18 * BN_X**2 + 6 * BN_X + 1 == R

G2_GEN = ((10857046999023057135944570762232829481370756359578518086990519993285655852781,
           11559732032986387107991004021392285783925812861821192530917403151452391805634),
          (8495653923123431417604973247489272438418190587263600148770280649306958101930,
           4082367875863433681332203403145435568316851327593401208105741076214120093531))


# ---- Fq2 -------------------------------------------------------------------
def f2_add(a, b):
    return ((a[0] + b[0]) % P, (a[1] + b[1]) % P)


def f2_sub(a, b):
    return ((a[0] - b[0]) % P, (a[1] - b[1]) % P)


def f2_neg(a):
    return (-a[0] % P, -a[1] % P)


def f2_mul(a, b):
    return ((a[0] * b[0] - a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)


def f2_sqr(a):
    return f2_mul(a, a)


def f2_scale(a, k):
    return (a[0] * k % P, a[1] * k % P)


def f2_conj(a):
    return (a[0], -a[1] % P)


def f2_inv(a):
    d = pow(a[0] * a[0] + a[1] * a[1], P - 2, P)
    return (a[0] * d % P, -a[1] * d % P)


def f2_pow(a, e):
    r = (1, 0)
    while e:
        if e & 1:
            r = f2_mul(r, a)
        a = f2_sqr(a)
        e >>= 1
    return r


F2_ZERO, F2_ONE = (0, 0), (1, 0)
TWIST_B = f2_mul((3, 0), f2_inv(XI))            # y^2 = x^3 + 3/xi on the twist

# Frobenius constants: gamma_k = xi^((p^k - 1)/6)
GAMMA1 = f2_pow(XI, (P - 1) // 6)
GAMMA2 = f2_pow(XI, (P * P - 1) // 6)
GAMMA3 = f2_pow(XI, (P**3 - 1) // 6)
assert GAMMA2[1] == 0 and f2_pow(XI, (P * P - 1) // 2) == (P - 1, 0)


# ---- G2 (affine on the twist; None = infinity) ---------------------------------
def g2_is_on_curve(Q):
    if Q is None:
        return True
    x, y = Q
    return f2_sqr(y) == f2_add(f2_mul(f2_sqr(x), x), TWIST_B)


def g2_neg(Q):
    return None if Q is None else (Q[0], f2_neg(Q[1]))


def g2_add(A, B):
    if A is None:
        return B
    if B is None:
        return A
    if A[0] == B[0]:
        if A[1] != B[1] or A[1] == F2_ZERO:
            return None
        lam = f2_mul(f2_scale(f2_sqr(A[0]), 3), f2_inv(f2_scale(A[1], 2)))
    else:
        lam = f2_mul(f2_sub(B[1], A[1]), f2_inv(f2_sub(B[0], A[0])))
    x3 = f2_sub(f2_sub(f2_sqr(lam), A[0]), B[0])
    return (x3, f2_sub(f2_mul(lam, f2_sub(A[0], x3)), A[1]))


def g2_mul(Q, k):
    k %= R
    acc = None
    while k:
        if k & 1:
            acc = g2_add(acc, Q)
        Q = g2_add(Q, Q)
        k >>= 1
    return acc


assert g2_is_on_curve(G2_GEN)


def g2_to_lem(Q):
    """128 bytes x.c0 | x.c1 | y.c0 | y.c1, 32-byte LE Montgomery each (the zkey header's X_2); infinity = zeros"""
    if Q is None:
        return bytes(128)
    return b"".join(((c << 256) % P).to_bytes(32, "little") for c in (Q[0][0], Q[0][1], Q[1][0], Q[1][1]))


def g2_from_lem(raw):
    if raw == bytes(128):
        return None
    rinv = pow(1 << 256, -1, P)
    c = [int.from_bytes(raw[i:i + 32], "little") * rinv % P for i in range(0, 128, 32)]
    return ((c[0], c[1]), (c[2], c[3]))


# ---- Fq12 = Fq2[w]/(w^6 - xi), six Fq2 coefficients ----------------------------
F12_ONE = [F2_ONE] + [F2_ZERO] * 5


def f12_mul(a, b):
    t = [F2_ZERO] * 11
    for i in range(6):
        if a[i] == F2_ZERO:
            continue
        for j in range(6):
            t[i + j] = f2_add(t[i + j], f2_mul(a[i], b[j]))
    return [f2_add(t[k], f2_mul(t[k + 6], XI)) if k < 5 else t[k] for k in range(6)]


def f12_sqr(a):
    return f12_mul(a, a)


def f12_conj(a):
    """a^(p^6): w -> -w"""
    return [a[k] if k % 2 == 0 else f2_neg(a[k]) for k in range(6)]


def f12_frob(a, k=1):
    """a^(p^k), k = 1, 2, 3: coefficient of w^m -> conj^k(coef) * gamma_k^m"""
    g = {1: GAMMA1, 2: GAMMA2, 3: GAMMA3}[k]
    out, gm = [], F2_ONE
    for m in range(6):
        c = a[m] if k == 2 else f2_conj(a[m])
        out.append(f2_mul(c, gm))
        gm = f2_mul(gm, g)
    return out


def _f6_inv(a0, a1, a2):
    """inverse in Fq6 = Fq2[v]/(v^3 - xi)"""
    t0 = f2_sub(f2_sqr(a0), f2_mul(XI, f2_mul(a1, a2)))
    t1 = f2_sub(f2_mul(XI, f2_sqr(a2)), f2_mul(a0, a1))
    t2 = f2_sub(f2_sqr(a1), f2_mul(a0, a2))
    d = f2_add(f2_mul(a0, t0), f2_mul(XI, f2_add(f2_mul(a2, t1), f2_mul(a1, t2))))
    di = f2_inv(d)
    return f2_mul(t0, di), f2_mul(t1, di), f2_mul(t2, di)


def f12_inv(a):
    """a^-1 = conj(a) / (a conj(a)); a conj(a) lies in Fq6 (even powers of w, v = w^2)"""
    c = f12_conj(a)
    n = f12_mul(a, c)
    assert n[1] == n[3] == n[5] == F2_ZERO
    i0, i1, i2 = _f6_inv(n[0], n[2], n[4])
    return f12_mul(c, [i0, F2_ZERO, i1, F2_ZERO, i2, F2_ZERO])


def f12_pow(a, e):
    r = F12_ONE
    while e:
        if e & 1:
            r = f12_mul(r, a)
        a = f12_sqr(a)
        e >>= 1
    return r


# ---- Miller loop ---------------------------------------------------------------
def _line(T, Q, Pt):
    """line through T and Q (tangent when equal) on the twist, evaluated at the G1 point Pt, and T + Q.
    Untwisting (x', y') -> (x' w^2, y' w^3) gives  l(P) = yP - lam xP w + (lam xT - yT) w^3."""
    if T[0] == Q[0] and T[1] == Q[1]:
        lam = f2_mul(f2_scale(f2_sqr(T[0]), 3), f2_inv(f2_scale(T[1], 2)))
    else:
        lam = f2_mul(f2_sub(Q[1], T[1]), f2_inv(f2_sub(Q[0], T[0])))
    x3 = f2_sub(f2_sub(f2_sqr(lam), T[0]), Q[0])
    y3 = f2_sub(f2_mul(lam, f2_sub(T[0], x3)), T[1])
    xp, yp = Pt
    l = [(yp, 0), f2_neg(f2_scale(lam, xp)), F2_ZERO, f2_sub(f2_mul(lam, T[0]), T[1]), F2_ZERO, F2_ZERO]
    return l, (x3, y3)


def g2_frobenius(Q):
    """pi(x', y') on the twist"""
    return (f2_mul(f2_conj(Q[0]), f2_sqr(GAMMA1)), f2_mul(f2_conj(Q[1]), f2_mul(f2_sqr(GAMMA1), GAMMA1)))


def miller_loop(Pt, Q):
    """f_{6x+2,Q}(P) * l_{T,pi(Q)}(P) * l_{T+pi(Q),-pi^2(Q)}(P); 1 when either point is infinity"""
    if Pt is None or Q is None:
        return list(F12_ONE)
    f = list(F12_ONE)
    T = Q
    for i in range(ATE_LOOP.bit_length() - 2, -1, -1):
        l, T2 = _line(T, T, Pt)
        f = f12_mul(f12_sqr(f), l)
        T = T2
        if (ATE_LOOP >> i) & 1:
            l, T = _line(T, Q, Pt)
            f = f12_mul(f, l)
    Q1 = g2_frobenius(Q)
    Q2n = (f2_mul(Q[0], f2_sqr(GAMMA2)), Q[1])  # -pi^2(Q): xi^((p^2-1)/2) = -1
    l, T = _line(T, Q1, Pt)
    f = f12_mul(f, l)
    l, T = _line(T, Q2n, Pt)
    return f12_mul(f, l)


def final_exp_easy(f):
    """f^((p^6 - 1)(p^2 + 1))"""
    f1 = f12_mul(f12_conj(f), f12_inv(f))
    return f12_mul(f12_frob(f1, 2), f1)


HARD_EXP = (P**4 - P * P + 1) // R


def final_exp_hard_chain(f):
    """f^((p^4 - p^2 + 1)/r) by the vectorial addition chain of Scott et al. (three powers by x, Frobenius maps);
    f must be unitary (after the easy part), so inversion is conjugation"""
    fx = f12_pow(f, BN_X)
    fx2 = f12_pow(fx, BN_X)
    fx3 = f12_pow(fx2, BN_X)
    y0 = f12_mul(f12_mul(f12_frob(f, 1), f12_frob(f, 2)), f12_frob(f, 3))
    y1 = f12_conj(f)
    y2 = f12_frob(fx2, 2)
    y3 = f12_conj(f12_frob(fx, 1))
    y4 = f12_conj(f12_mul(fx, f12_frob(fx2, 1)))
    y5 = f12_conj(fx2)
    y6 = f12_conj(f12_mul(fx3, f12_frob(fx3, 1)))
    t0 = f12_mul(f12_mul(f12_sqr(y6), y4), y5)
    t1 = f12_mul(f12_mul(y3, y5), t0)
    t0 = f12_mul(t0, y2)
    t1 = f12_sqr(f12_mul(f12_sqr(t1), t0))
    t0 = f12_mul(t1, y1)
    t1 = f12_mul(t1, y0)
    return f12_mul(t1, f12_sqr(t0))


def final_exp(f):
    return f12_pow(final_exp_easy(f), HARD_EXP)


def pairing(Pt, Q):
    """e(P, Q) in GT, P in G1 (affine ints or None), Q in G2"""
    return final_exp(miller_loop(Pt, Q))


def pairing_eq(pairs):
    """prod e(P_i, Q_i) == 1  (curve.pairingEq)"""
    f = list(F12_ONE)
    for Pt, Q in pairs:
        f = f12_mul(f, miller_loop(Pt, Q))
    return final_exp(f) == F12_ONE
