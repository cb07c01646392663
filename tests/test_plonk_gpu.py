"""Prover parity (config 2 at oracle-sized domains): GPU proof bytes == Python
oracle proof bytes under injected blinders, accepted by the independently
written verifier; snarkjs' error behaviour on bad witnesses."""
import random

import pytest

from oracle import bn254 as b
from oracle import plonk as oplonk
from oracle.binfile import write_wtns
from oracle.keccak import hash_to_fr
from tests.circuits_util import random_circuit

pytestmark = pytest.mark.gpu

TAU = hash_to_fr(b"nzcb-b200-tau")
_srs_cache = {}


def srs(count):
    key = max(count, 70)
    if key not in _srs_cache:
        _srs_cache[key] = b.srs_g1(TAU, key)
    return _srs_cache[key][:count]


def _setup(seed, **kw):
    r, w = random_circuit(seed, **kw)
    gates, _, _ = oplonk.r1cs_to_plonk(r)
    n = 1 << max(3, (len(gates) - 1).bit_length())
    return r, w, oplonk.setup(r, srs(n + 6))


@pytest.mark.parametrize("seed,kw", [
    (1, dict(n_out=1, n_in=2, n_mul=3)),                 # domain 8/16
    (2, dict(n_out=2, n_in=3, n_mul=20)),                # 64
    (3, dict(n_out=3, n_in=4, n_mul=60, public_inputs=2)),  # 5 public signals
    (4, dict(n_out=0, n_in=3, n_mul=25)),                # nPublic = 0
])
def test_proof_bytes_match_oracle(ctx, seed, kw):
    from nzcb_circom_b200.snarkjs import ZKey, plonk

    r, w, zkey = _setup(seed, **kw)
    rng = random.Random(seed)
    blinders = [rng.randrange(b.R_MOD) for _ in range(9)]
    wt = write_wtns(w)
    exp_proof, exp_pub, _ = oplonk.prove(zkey, wt, blinders)
    zk = ZKey(zkey, ctx)
    got, pub = plonk.prove(zk, wt, blinders=blinders, raw=True)
    assert got == oplonk.proof_to_bytes(exp_proof)
    assert pub == [str(x) for x in exp_pub]
    assert plonk.proof_json(got, ctx) == oplonk.proof_to_json(exp_proof)
    vk = oplonk.verification_key(zkey)
    assert oplonk.verify_with_trapdoor(vk, [int(x) for x in pub], oplonk.proof_from_bytes(got), TAU)
    # random blinders: different proof, still verifies
    got2, pub2 = plonk.prove(zk, wt, raw=True)
    assert got2 != got and pub2 == pub
    assert oplonk.verify_with_trapdoor(vk, [int(x) for x in pub2], oplonk.proof_from_bytes(got2), TAU)
    zk.close()


def test_bad_witness_errors(ctx):
    from nzcb_circom_b200 import NzcbError
    from nzcb_circom_b200.snarkjs import ZKey, plonk

    r, w, zkey = _setup(2, n_out=2, n_in=3, n_mul=20)
    zk = ZKey(zkey, ctx)
    bl = list(range(1, 10))
    with pytest.raises(NzcbError, match="Invalid witness length"):
        plonk.prove(zk, write_wtns(w[:-1]), blinders=bl)
    bad = list(w)
    bad[len(w) // 2] = (bad[len(w) // 2] + 1) % b.R_MOD
    with pytest.raises(NzcbError) as ei:
        plonk.prove(zk, write_wtns(bad), blinders=bl)
    with pytest.raises(oplonk.ProverError) as eo:
        oplonk.prove(zkey, write_wtns(bad), bl)
    assert str(eo.value) in str(ei.value)  # same snarkjs message
    # a failed proof does not poison the context
    got, _ = plonk.prove(zk, write_wtns(w), blinders=bl, raw=True)
    assert got == oplonk.proof_to_bytes(oplonk.prove(zkey, write_wtns(w), bl)[0])
    # batch: per-proof status, good proofs unaffected
    res = plonk.prove_batch(zk, [write_wtns(w), write_wtns(bad), write_wtns(w)], [bl, bl, bl])
    assert [s for _, _, s in res] == [0, res[1][2], 0] and res[1][2] < 0
    assert res[0][0] == got and res[2][0] == got and res[1][0] is None
    zk.close()
