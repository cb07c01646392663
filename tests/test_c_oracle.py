"""Pins the C restatement (oracle/c/nzcb_oracle.c, the cpu_baseline / --impl reference engine)
against the Python big-int oracle: Keccak, NTT, MSM, witness program, prover -- byte for byte."""
import os
import random

import pytest

from oracle import bn254 as b
from oracle import c_oracle as C
from oracle import ntt as ontt
from oracle import plonk as oplonk
from oracle import witness_vm as vm
from oracle.binfile import write_wtns
from oracle.keccak import hash_to_fr, keccak256
from tests.circuits_util import random_circuit

TAU = hash_to_fr(b"nzcb-b200-tau")


def test_keccak():
    for n in (0, 1, 135, 136, 137, 288, 1000):
        d = os.urandom(n)
        assert C.keccak256(d) == keccak256(d)


@pytest.mark.parametrize("log_n", [1, 3, 8, 11])
def test_ntt(log_n):
    rng = random.Random(log_n)
    v = [rng.randrange(b.R_MOD) for _ in range(1 << log_n)]
    buf = b"".join(b.to_lem(x) for x in v)
    assert C.ntt(buf) == b"".join(b.to_lem(x) for x in ontt.fft(v))
    assert C.ntt(buf, True) == b"".join(b.to_lem(x) for x in ontt.ifft(v))


@pytest.mark.parametrize("n", [1, 2, 50, 700])
def test_msm(n):
    rng = random.Random(n)
    P = b.g1_mul(b.G1_GEN, rng.randrange(1, b.R_MOD))
    step = b.g1_mul(b.G1_GEN, 7)
    pts = []
    for _ in range(n):
        pts.append(P)
        P = b.g1_add(P, step)
    sc = [rng.randrange(b.R_MOD) for _ in range(n)]
    if n > 10:
        sc[0], sc[1], pts[2] = 0, b.R_MOD - 1, None
    got = C.msm(b"".join(b.g1_to_lem(p) for p in pts), b"".join(b.to_le(s) for s in sc))
    assert b.g1_from_lem(got) == b.g1_msm(pts, sc)


def test_witness_program():
    from nzcb_circom_b200.circom_tester import compile_circuit

    art = compile_circuit("skipValue5_test")
    prog = vm.Program(art.wprog_bytes())
    for inp in ({"bytes": [0x83, 23, 23, 23, 0], "pos": 0}, {"bytes": [0x61, 0x61, 0, 0, 0], "pos": 0}):
        flat = art.flatten_input(inp)
        rc, raw = C.witness(art.wprog_bytes(), b"".join(v.to_bytes(32, "little") for v in flat), art.n_total)
        exp = vm.run(prog, flat)
        assert rc == 0 and [int.from_bytes(raw[i * 32:(i + 1) * 32], "little") for i in range(art.n_witness)] == exp
    art = compile_circuit("quinSelector3_test")
    flat = art.flatten_input({"in": [1, 2, 3], "index": 3})
    rc, _ = C.witness(art.wprog_bytes(), b"".join(v.to_bytes(32, "little") for v in flat), art.n_total)
    assert rc == -6  # Assert Failed


@pytest.mark.parametrize("seed,kw", [(1, dict(n_out=1, n_in=2, n_mul=3)), (3, dict(n_out=3, n_in=4, n_mul=60, public_inputs=2)),
                                     (4, dict(n_out=0, n_in=3, n_mul=25))])
def test_prover(seed, kw):
    r, w = random_circuit(seed, **kw)
    gates, _, _ = oplonk.r1cs_to_plonk(r)
    n = 1 << max(3, (len(gates) - 1).bit_length())
    zkey = oplonk.setup(r, b.srs_g1(TAU, n + 6))
    rng = random.Random(seed)
    bl = [rng.randrange(b.R_MOD) for _ in range(9)]
    exp, pub, _ = oplonk.prove(zkey, write_wtns(w), bl)
    rc, proof, cpub = C.prove(zkey, write_wtns(w), bl, r.n_public)
    assert rc == 0 and proof == oplonk.proof_to_bytes(exp) and cpub == pub
    bad = list(w)
    bad[len(w) // 2] ^= 1
    assert C.prove(zkey, write_wtns(bad), bl, r.n_public)[0] in (-4, -5)
    assert C.prove(zkey, write_wtns(w[:-1]), bl, r.n_public)[0] == -3
