"""`snarkjs powersoftau new` / `plonk setup` roles on the GPU vs. the oracle:
SRS points and the whole zkey file must be byte-identical."""
import pytest

from oracle import bn254 as b
from oracle import plonk as oplonk
from oracle.binfile import write_r1cs
from oracle.keccak import hash_to_fr
from tests.circuits_util import random_circuit

pytestmark = pytest.mark.gpu
TAU = hash_to_fr(b"nzcb-b200-tau")


def test_srs_matches_oracle(ctx):
    from nzcb_circom_b200.snarkjs import powersoftau

    raw = powersoftau.new_g1(TAU, 300, ctx)
    exp = b.srs_g1(TAU, 40)
    for i, P in enumerate(exp):
        assert b.g1_from_lem(raw[i * 64:(i + 1) * 64]) == P
    for i in (77, 299):
        assert b.g1_from_lem(raw[i * 64:(i + 1) * 64]) == b.g1_mul(b.G1_GEN, pow(TAU, i, b.R_MOD))


@pytest.mark.parametrize("seed,kw", [
    (1, dict(n_out=1, n_in=2, n_mul=3)),
    (3, dict(n_out=3, n_in=4, n_mul=60, public_inputs=2)),
    (4, dict(n_out=0, n_in=3, n_mul=25)),
])
def test_zkey_bytes_match_oracle(ctx, seed, kw):
    from nzcb_circom_b200.snarkjs import plonk, powersoftau

    r, _ = random_circuit(seed, **kw)
    gates, _, _ = oplonk.r1cs_to_plonk(r)
    n = 1 << max(3, (len(gates) - 1).bit_length())
    srs_raw = powersoftau.new_g1(TAU, n + 6, ctx)
    srs = [b.g1_from_lem(srs_raw[i * 64:(i + 1) * 64]) for i in range(n + 6)]
    x2 = bytes(range(128))
    exp = oplonk.setup(r, srs, x2)
    got = plonk.setup(write_r1cs(r), srs_raw, x2, ctx)
    assert len(got) == len(exp)
    assert got == exp
