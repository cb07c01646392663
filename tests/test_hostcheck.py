"""The product's field / curve headers (csrc/fp.cuh, g1.cuh, keccak.h) compiled for the host and checked against the
Python big-int oracle: the exact arithmetic the kernels run (portable CIOS multiply; the PTX carry-chain version is
checked against it on the device by nzcb_selftest_mul), without a GPU."""
import ctypes
import os
import random
import subprocess

import pytest

from oracle import bn254 as b
from oracle.keccak import keccak256

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostcheck")
SRC = os.path.join(HERE, "fieldcheck.cpp")
LIB = os.path.join(HERE, "libfieldcheck.so")


@pytest.fixture(scope="module")
def fc():
    deps = [SRC] + [os.path.join(HERE, "..", "..", "nzcb_circom_b200", "csrc", f) for f in ("fp.cuh", "g1.cuh", "keccak.h")]
    if not os.path.exists(LIB) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in deps):
        subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-x", "c++", SRC, "-o", LIB], check=True)
    return ctypes.CDLL(LIB)


def _field(fc, field, op, a, bb=0):
    mod = b.R_MOD if field == 0 else b.P_MOD
    out = ctypes.create_string_buffer(32)
    assert fc.fc_field_op(field, op, (a % (1 << 256)).to_bytes(32, "little"), (bb % (1 << 256)).to_bytes(32, "little"), out) == 0
    return int.from_bytes(out.raw, "little"), mod


@pytest.mark.parametrize("field", [0, 1])
def test_field_ops(fc, field):
    mod = b.R_MOD if field == 0 else b.P_MOD
    rng = random.Random(field)
    rinv = pow(1 << 256, -1, mod)
    vals = [0, 1, mod - 1, mod - 2, 2, (1 << 253) % mod] + [rng.randrange(mod) for _ in range(40)]
    for x in vals:
        for y in vals[:12]:
            assert _field(fc, field, 0, x, y)[0] == (x + y) % mod
            assert _field(fc, field, 1, x, y)[0] == (x - y) % mod
            assert _field(fc, field, 2, x, y)[0] == x * y * rinv % mod  # Montgomery product
        assert _field(fc, field, 4, x)[0] == x * (1 << 256) % mod
        assert _field(fc, field, 5, x)[0] == x * rinv % mod
        assert _field(fc, field, 6, x)[0] == (-x) % mod
    for x in vals[1:12]:  # Montgomery in / out inverse
        xm = x * (1 << 256) % mod
        assert _field(fc, field, 3, xm)[0] == pow(x, -1, mod) * (1 << 256) % mod
    assert _field(fc, field, 3, 0)[0] == 0


def test_g1_ops(fc):
    rng = random.Random(7)
    P = b.g1_mul(b.G1_GEN, rng.randrange(1, b.R_MOD))
    Q = b.g1_mul(b.G1_GEN, rng.randrange(1, b.R_MOD))

    def op(code, A, B, k=0):
        out = ctypes.create_string_buffer(64)
        assert fc.fc_g1_op(code, b.g1_to_lem(A), b.g1_to_lem(B), ctypes.c_uint64(k), out) == 0
        return b.g1_from_lem(out.raw)

    assert op(0, P, Q) == b.g1_add(P, Q)
    assert op(0, P, P) == b.g1_add(P, P)          # the doubling branch of the mixed addition
    assert op(0, P, b.g1_neg(P)) is None          # P + (-P) = infinity
    assert op(0, P, None) == P and op(0, None, Q) == Q
    assert op(1, P, Q) == b.g1_add(P, Q)          # full XYZZ + XYZZ with non-trivial ZZ on both sides
    assert op(2, P, P) == b.g1_add(P, P)
    for k in (0, 1, 2, 3, 255, 0xDEADBEEFCAFE):
        assert op(3, P, P, k) == (b.g1_mul(P, k) if k else None)


def test_keccak(fc):
    for n in (0, 1, 135, 136, 137, 288, 1000):
        d = os.urandom(n)
        out = ctypes.create_string_buffer(32)
        fc.fc_keccak256(d, ctypes.c_size_t(n), out)
        assert out.raw == keccak256(d)


def test_affine_rounds(fc):
    """msm_affine.cuh: R halving rounds over a padded, bucket-sorted reference list leave, per bucket, points whose
    sum is the bucket's sum -- with null padding, infinity bases, negated references, equal points (doubling) and
    cancelling pairs inside the batches, and batch / chunk sizes that do not divide the list."""
    rng = random.Random(11)
    n_bases = 90
    pts = []
    P = b.g1_mul(b.G1_GEN, rng.randrange(1, b.R_MOD))
    step = b.g1_mul(b.G1_GEN, rng.randrange(1, b.R_MOD))
    for _ in range(n_bases):
        pts.append(P)
        P = b.g1_add(P, step)
    pts[5] = None                     # infinity base
    pts[7] = pts[6]                   # equal bases
    pts[9] = b.g1_neg(pts[8])         # cancelling pair
    bases = b"".join(b.g1_to_lem(p) for p in pts)
    NULL = 0xFFFFFFFF
    for R, n_buckets in ((1, 40), (3, 300), (3, 7), (2, 1100)):
        refs, want = [], []
        for bk in range(n_buckets):
            cnt = rng.choice([0, 1, 2, 3, 5, 8, 9, 17, 40]) if n_buckets < 1000 else rng.choice([0, 1, 2, 7])
            seg = []
            for _ in range(cnt):
                i = rng.randrange(n_bases)
                if rng.random() < 0.3:
                    i = rng.choice([5, 6, 7, 8, 9])  # provoke inf + P, P + P, P - P
                seg.append(i | (rng.getrandbits(1) << 31))
            if bk == 0 and R == 3:
                seg = [6, 7, 6 | 1 << 31, 7 | 1 << 31, 8, 9, 5, 5, 8 | 1 << 31, 9 | 1 << 31][:cnt] + seg[10:]
            acc = None
            for e in seg:
                p = pts[e & 0x7FFFFFFF]
                acc = b.g1_add(acc, b.g1_neg(p) if (e >> 31) and p is not None else p)
            pad = (-len(seg)) % (1 << R)
            want.append((len(refs) >> R, (len(seg) + pad) >> R, acc))
            refs += seg + [NULL] * pad
        n_refs = len(refs)
        arr = (ctypes.c_uint32 * max(1, n_refs))(*refs)
        out = ctypes.create_string_buffer(max(64, (n_refs >> R) * 64))
        assert fc.fc_affine_rounds(bases, n_bases, arr, n_refs, R, out) == 0
        for lo, cnt, acc in want:
            got = None
            for k in range(lo, lo + cnt):
                got = b.g1_add(got, b.g1_from_lem(out.raw[64 * k:64 * k + 64]))
            assert got == acc
