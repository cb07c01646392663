"""`.ptau` reader: `snarkjs plonk setup circuit.r1cs pot.ptau circuit.zkey` with the ceremony file
itself.  The file comes from the oracle's writer (oracle/ptau.py, a small known-trapdoor ceremony); the product
parses it (header, tauG1, tauG2[1], the "prepared" check) and must emit the same zkey bytes as the oracle's setup
over the same points."""
import pytest

from oracle import bn254 as b
from oracle import pairing as pg
from oracle import plonk as oplonk
from oracle import ptau as optau
from oracle.binfile import read_binfile, section, write_r1cs
from oracle.keccak import hash_to_fr
from tests.circuits_util import random_circuit

TAU = hash_to_fr(b"nzcb-b200-tau")
POWER = 6


@pytest.fixture(scope="module")
def ptau_file():
    return optau.write_ptau(TAU, POWER)


def test_ptau_layout_and_info(ptau_file):
    from nzcb_circom_b200.snarkjs import powersoftau

    info = powersoftau.info(ptau_file)
    assert info == {"power": POWER, "ceremonyPower": POWER, "nTauG1": 2 * (1 << POWER) - 1, "prepared": True}
    _, secs = read_binfile(ptau_file, b"ptau")
    s2, s3, s12 = section(ptau_file, secs, 2), section(ptau_file, secs, 3), section(ptau_file, secs, 12)
    assert b.g1_from_lem(s2[:64]) == b.G1_GEN and b.g1_from_lem(s2[64:128]) == b.g1_mul(b.G1_GEN, TAU)
    assert pg.g2_from_lem(s3[128:256]) == pg.g2_mul(pg.G2_GEN, TAU)
    assert len(s12) == 64 * (2 * (1 << POWER) - 1)
    # the Lagrange points of the top domain commit to the same polynomial as the monomial points: sum_i L_i(tau) = 1
    top = s12[64 * ((1 << POWER) - 1):]
    acc = None
    for i in range(1 << POWER):
        acc = b.g1_add(acc, b.g1_from_lem(top[64 * i:64 * i + 64]))
    assert acc == b.G1_GEN
    assert powersoftau.info(optau.write_ptau(TAU, 2, prepared=False))["prepared"] is False
    for bad in (b"ptax" + ptau_file[4:], ptau_file[:40], ptau_file[:-50]):
        with pytest.raises(ValueError):
            powersoftau.info(bad)


@pytest.mark.gpu
def test_setup_from_ptau_matches_the_oracle(ctx, ptau_file):
    from nzcb_circom_b200 import NzcbError
    from nzcb_circom_b200.snarkjs import plonk

    r, _ = random_circuit(3, n_out=2, n_in=4, n_mul=8, public_inputs=2)
    gates, _, _ = oplonk.r1cs_to_plonk(r)
    n = 1 << max(3, (len(gates) - 1).bit_length())
    assert n <= 1 << POWER
    srs = b.srs_g1(TAU, n + 6)
    exp = oplonk.setup(r, srs, pg.g2_to_lem(pg.g2_mul(pg.G2_GEN, TAU)))
    got = plonk.setup_ptau(write_r1cs(r), ptau_file, ctx)
    assert bytes(got) == bytes(exp)
    # snarkjs' two refusals
    with pytest.raises(NzcbError, match="not prepared"):
        plonk.setup_ptau(write_r1cs(r), optau.write_ptau(TAU, POWER, prepared=False), ctx)
    big, _ = random_circuit(5, n_out=1, n_in=3, n_mul=30)
    with pytest.raises(NzcbError, match="circuit too big"):
        plonk.setup_ptau(write_r1cs(big), optau.write_ptau(TAU, 3), ctx)
