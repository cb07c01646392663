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
