"""Config 3 (primitive sweep) parity: BN254 Fr NTT and G1 MSM through the C ABI
vs. the Python big-int oracle (bit-exact), plus size-independent algebraic
properties at the BASELINE.json sizes.  SURVEY.md section 4 (v)."""
import random

import pytest

from oracle import bn254 as b
from oracle import ntt as ontt

pytestmark = pytest.mark.gpu


def _fr_buf(vals):
    return b"".join(b.to_lem(v) for v in vals)


def _fr_unbuf(buf):
    return [b.from_lem(buf[i:i + 32]) for i in range(0, len(buf), 32)]


@pytest.mark.parametrize("log_n", [1, 2, 3, 4, 5, 6, 7, 9, 10, 11, 12, 13, 14, 15])
def test_ntt_matches_oracle(ctx, log_n):
    from nzcb_circom_b200.ffjavascript import Fr

    rng = random.Random(0x6E7A6362 + log_n)
    vals = [rng.randrange(b.R_MOD) for _ in range(1 << log_n)]
    vals[0] = 0
    vals[-1] = b.R_MOD - 1
    got = _fr_unbuf(Fr.fft(_fr_buf(vals), ctx))
    assert got == ontt.fft(vals)
    got_i = _fr_unbuf(Fr.ifft(_fr_buf(vals), ctx))
    assert got_i == ontt.ifft(vals)


@pytest.mark.parametrize("log_n", [16, 17, 18, 19, 20, 21, 22, 23])
def test_ntt_roundtrip_and_pointcheck_large(ctx, log_n):
    """iNTT(NTT(x)) == x bit-exact, and X[k] == sum_j x_j w^(jk) for a sparse x."""
    import numpy as np
    from nzcb_circom_b200.ffjavascript import Fr

    n = 1 << log_n
    rng = np.random.default_rng(log_n)
    raw = rng.integers(0, 256, size=n * 32, dtype=np.uint8)
    raw.reshape(n, 32)[:, 31] &= 0x1F  # < 2^253 < r: valid (arbitrary) Montgomery residues
    buf = raw.tobytes()
    fwd = Fr.fft(buf, ctx)
    assert Fr.ifft(fwd, ctx) == buf
    # sparse input: three non-zero coefficients -> closed form at a few output points
    idx = [0, 5, n - 3]
    coef = [7, b.R_MOD - 2, 123456789]
    sp = bytearray(n * 32)
    for i, c in zip(idx, coef):
        sp[i * 32:(i + 1) * 32] = b.to_lem(c)
    out = Fr.fft(bytes(sp), ctx)
    w = b.fr_root(log_n)
    for k in [0, 1, 2, n // 2 + 1, n - 1, 12345 % n]:
        exp = sum(c * pow(w, i * k, b.R_MOD) for i, c in zip(idx, coef)) % b.R_MOD
        assert b.from_lem(out[k * 32:(k + 1) * 32]) == exp


def _points(rng, n):
    # distinct multiples of G, built incrementally (cheap in Python)
    pts = []
    P = b.g1_mul(b.G1_GEN, rng.randrange(1, b.R_MOD))
    step = b.g1_mul(b.G1_GEN, rng.randrange(1, b.R_MOD))
    for _ in range(n):
        pts.append(P)
        P = b.g1_add(P, step)
    return pts


@pytest.mark.parametrize("n", [0, 1, 2, 3, 31, 33, 100, 1000, 5000])
def test_msm_matches_oracle(ctx, n):
    from nzcb_circom_b200.ffjavascript import G1

    rng = random.Random(n + 1)
    pts = _points(rng, n)
    sc = [rng.randrange(b.R_MOD) for _ in range(n)]
    if n >= 3:
        sc[0] = 0
        sc[1] = b.R_MOD - 1
        sc[2] = 1
    if n >= 33:
        pts[5] = None  # infinity base (all-zero bytes)
        pts[7] = pts[6]  # equal bases -> doubling path when scalars collide
        sc[7] = sc[6]
        pts[9] = b.g1_neg(pts[8])  # cancelling pair
        sc[9] = sc[8]
    bases = b"".join(b.g1_to_lem(P) for P in pts)
    scal = b"".join(b.to_le(s) for s in sc)
    got = b.g1_from_lem(G1.multiExpAffine(bases, scal, ctx))
    assert got == b.g1_msm(pts, sc)


def test_msm_small_scalars(ctx):
    """0/1/byte scalars (what wire values look like): heavy bucket imbalance."""
    from nzcb_circom_b200.ffjavascript import G1

    rng = random.Random(99)
    n = 4096
    pts = _points(rng, n)
    sc = [rng.choice([0, 1, 1, 1, rng.randrange(256)]) for _ in range(n)]
    bases = b"".join(b.g1_to_lem(P) for P in pts)
    scal = b"".join(b.to_le(s) for s in sc)
    got = b.g1_from_lem(G1.multiExpAffine(bases, scal, ctx))
    assert got == b.g1_msm(pts, sc)


# ---- fixed-base table mode (how the prover runs its nine MSMs) ------------------------------
@pytest.mark.parametrize("n", [1, 2, 33, 100, 1000, 5000])
def test_table_msm_matches_oracle(ctx, n):
    from nzcb_circom_b200.ffjavascript import G1Table

    rng = random.Random(n + 7)
    pts = _points(rng, n)
    if n >= 33:
        pts[5] = None
        pts[7] = pts[6]
        pts[9] = b.g1_neg(pts[8])
    tab = G1Table(b"".join(b.g1_to_lem(P) for P in pts), ctx)
    sc = [rng.randrange(b.R_MOD) for _ in range(n)]
    if n >= 33:
        sc[0], sc[1], sc[2], sc[3] = 0, b.R_MOD - 1, 1, (b.R_MOD - 1) // 2
        sc[4] = (b.R_MOD + 1) // 2
        sc[7] = sc[6]
        sc[9] = sc[8]
    assert b.g1_from_lem(tab.multiExpAffine(b"".join(b.to_le(s) for s in sc))) == b.g1_msm(pts, sc)
    # a prefix of the bases, and a batch of jobs of different lengths (A/B/C, T1/T2/T3 style)
    m = max(1, n // 2)
    jobs = [sc, [rng.randrange(b.R_MOD) for _ in range(m)], [rng.choice([0, 1, b.R_MOD - 1, 255]) for _ in range(n)]]
    got = tab.batch([b"".join(b.to_le(s) for s in j) for j in jobs])
    for j, g in zip(jobs, got):
        assert b.g1_from_lem(g) == b.g1_msm(pts[:len(j)], j)
    tab.close()


def test_table_msm_giant_buckets(ctx):
    """all scalars equal: one bucket per window holds every point (the segmented reduction's worst case)"""
    from nzcb_circom_b200.ffjavascript import G1Table

    rng = random.Random(5)
    n = 3000
    pts = _points(rng, n)
    tab = G1Table(b"".join(b.g1_to_lem(P) for P in pts), ctx)
    S = None
    for P in pts:
        S = b.g1_add(S, P)
    for s in (1, b.R_MOD - 1, 0x1234567, rng.randrange(b.R_MOD)):
        got = b.g1_from_lem(tab.multiExpAffine(b.to_le(s) * n))
        assert got == b.g1_mul(S, s)
    tab.close()


@pytest.mark.parametrize("log_n", [16, 17, 18, 19, 20, 21, 22])
def test_table_msm_srs_identity_large(ctx, log_n):
    """sum_i c_i [tau^i]G == p(tau) G for the synthetic SRS, full-size scalars and wire-like small ones;
    table mode and one-shot window mode agree bit for bit."""
    import numpy as np
    from nzcb_circom_b200.ffjavascript import G1, G1Table
    from nzcb_circom_b200.snarkjs import powersoftau
    from oracle.keccak import hash_to_fr

    tau = hash_to_fr(b"nzcb-b200-tau")
    n = (1 << log_n) + 6
    srs = powersoftau.new_g1(tau, n, ctx)
    tab = G1Table(srs, ctx)
    rng = np.random.default_rng(log_n)
    raw = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    raw[:, 31] &= 0x1F  # < 2^253 < r
    small = np.zeros((n, 32), dtype=np.uint8)
    small[:, 0] = rng.integers(0, 2, size=n, dtype=np.uint8)
    small[::7, 0] = rng.integers(0, 256, size=len(small[::7]), dtype=np.uint8)
    for arr in (raw, small):
        buf = arr.tobytes()
        # p(tau) by Horner over numpy-decoded ints would be slow in Python at 2^21: use chunks of 8 bytes
        coef = [int.from_bytes(buf[i * 32:(i + 1) * 32], "little") for i in range(n)]
        acc = 0
        for c in reversed(coef):
            acc = (acc * tau + c) % b.R_MOD
        exp = b.g1_mul(b.G1_GEN, acc)
        got_t = tab.multiExpAffine(buf)
        assert b.g1_from_lem(got_t) == exp
        assert G1.multiExpAffine(srs, buf, ctx) == got_t
    tab.close()


@pytest.mark.parametrize("log_n", [0, 1, 2, 5, 8])
def test_lagrange_basis_matches_trapdoor(ctx, log_n):
    """group inverse DFT of the SRS: out[i] == L_i(tau) * G with L_i(tau) = (1/n) sum_j w^(-ij) tau^j"""
    from nzcb_circom_b200.snarkjs import powersoftau
    from oracle.keccak import hash_to_fr

    tau = hash_to_fr(b"nzcb-b200-tau")
    n = 1 << log_n
    srs = powersoftau.new_g1(tau, n, ctx)
    got = powersoftau.lagrange_g1(srs, log_n, ctx)
    w = b.fr_root(log_n) if log_n else 1
    n_inv = pow(n, -1, b.R_MOD)
    for i in sorted({0, 1 % n, n // 2, n - 1}):
        wi = pow(w, -i, b.R_MOD)
        li = sum(pow(wi, j, b.R_MOD) * pow(tau, j, b.R_MOD) for j in range(n)) * n_inv % b.R_MOD
        assert b.g1_from_lem(got[64 * i:64 * i + 64]) == b.g1_mul(b.G1_GEN, li)
