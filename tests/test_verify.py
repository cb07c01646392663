"""plonk.verify + BN254 pairing.

CPU: the pairing oracle (oracle/pairing.py) against its defining properties, the oracle verifier against the
known-trapdoor form, and the device headers (csrc/pairing.cuh, verify.cuh, keccak_hd.cuh) compiled for the host
against the oracle -- GT values bit for bit.  GPU: nzcb_pairing_eq / nzcb_plonk_verify_batch / nzcb_srs_g2 through
the C ABI against the oracle."""
import base64
import ctypes
import json
import os
import random
import struct
import subprocess

import pytest

from oracle import bn254 as b
from oracle import pairing as pg
from oracle import plonk as op
from oracle.keccak import hash_to_fr, keccak256

HERE = os.path.dirname(os.path.abspath(__file__))
TAU = hash_to_fr(b"nzcb-b200-tau")
X2_LEM = pg.g2_to_lem(pg.g2_mul(pg.G2_GEN, TAU))
FIXTURES = ["tiny", "small", "nopublic"]


def f12_lem(f):
    return b"".join(((c << 256) % b.P_MOD).to_bytes(32, "little") for co in f for c in co)


def _fixture(name):
    with open(os.path.join(HERE, "golden", f"plonk_{name}.json")) as fh:
        fx = json.load(fh)
    zkey = base64.b64decode(fx["zkey_b64"])
    vk = op.verification_key(zkey)
    vk["X_2"] = X2_LEM  # the fixtures' zkeys carry no X_2; the SRS trapdoor is TAU
    return fx, zkey, vk, bytes.fromhex(fx["proof_hex"]), [int(x) for x in fx["public_signals"]]


def _mutations(proof, pub, rng):
    """[(label, proof bytes, publics, expected)] around one valid proof"""
    out = [("valid", proof, pub, True)]
    for off, label in ((600, "eval_a"), (799, "eval_r"), (10, "A.x"), (200, "Z"), (500, "Wxi")):
        bad = bytearray(proof)
        bad[off] ^= 1
        out.append((label, bytes(bad), pub, False))
    # A replaced by another curve point: well formed, fails only at the pairing
    other = b.g1_mul(b.G1_GEN, rng.randrange(1, b.R_MOD))
    out.append(("A_other_point", b.g1_to_be(other) + proof[64:], pub, False))
    out.append(("A_infinity", bytes(64) + proof[64:], pub, False))
    big = bytearray(proof)
    big[576:608] = (b.R_MOD + 5).to_bytes(32, "big")  # eval_a >= r
    out.append(("eval_not_reduced", bytes(big), pub, False))
    bigx = bytearray(proof)
    bigx[0:32] = (b.P_MOD + 1).to_bytes(32, "big")  # x >= p
    out.append(("coordinate_not_reduced", bytes(bigx), pub, False))
    if pub:
        out.append(("public_changed", proof, [pub[0] ^ 1] + pub[1:], False))
        out.append(("public_plus_r", proof, [pub[0] + b.R_MOD] + pub[1:], pub[0] + b.R_MOD < (1 << 256)))
    return out


# ---------------------------------------------------------------- oracle
def test_pairing_oracle_properties():
    rng = random.Random(1)
    e1 = pg.pairing(b.G1_GEN, pg.G2_GEN)
    assert e1 != pg.F12_ONE and pg.f12_pow(e1, b.R_MOD) == pg.F12_ONE
    a, c = rng.randrange(1, b.R_MOD), rng.randrange(1, b.R_MOD)
    assert pg.pairing(b.g1_mul(b.G1_GEN, a), pg.g2_mul(pg.G2_GEN, c)) == pg.f12_pow(e1, a * c % b.R_MOD)
    assert pg.pairing(None, pg.G2_GEN) == pg.F12_ONE and pg.pairing(b.G1_GEN, None) == pg.F12_ONE
    assert pg.g2_mul(pg.G2_GEN, b.R_MOD) is None and pg.g2_is_on_curve(pg.g2_mul(pg.G2_GEN, a))
    # the addition chain of the hard part is the plain power
    f = pg.final_exp_easy(pg.miller_loop(b.g1_mul(b.G1_GEN, a), pg.G2_GEN))
    assert pg.final_exp_hard_chain(f) == pg.f12_pow(f, pg.HARD_EXP)
    assert pg.pairing_eq([(b.g1_mul(b.G1_GEN, a), pg.G2_GEN), (b.g1_neg(b.G1_GEN), pg.g2_mul(pg.G2_GEN, a))])
    assert not pg.pairing_eq([(b.g1_mul(b.G1_GEN, a), pg.G2_GEN), (b.G1_GEN, pg.g2_mul(pg.G2_GEN, a))])


@pytest.mark.parametrize("name", FIXTURES)
def test_oracle_verify_agrees_with_the_trapdoor_form(name):
    fx, zkey, vk, proof, pub = _fixture(name)
    for label, pr, pu, expected in _mutations(proof, pub, random.Random(2)):
        p = op.proof_from_bytes(pr)
        assert op.verify(vk, pu, p) == expected, label
        assert op.verify_with_trapdoor(vk, pu, p, TAU) == expected, label
    wrong = dict(vk)
    wrong["X_2"] = pg.g2_to_lem(pg.g2_mul(pg.G2_GEN, TAU + 1))
    assert not op.verify(wrong, pub, op.proof_from_bytes(proof))


# ---------------------------------------------------------------- the device headers on the host
@pytest.fixture(scope="module")
def fc():
    hc = os.path.join(HERE, "hostcheck")
    src, lib = os.path.join(hc, "fieldcheck.cpp"), os.path.join(hc, "libfieldcheck.so")
    csrc = os.path.join(HERE, "..", "nzcb_circom_b200", "csrc")
    deps = [src] + [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith((".cuh", ".h"))]
    if not os.path.exists(lib) or any(os.path.getmtime(d) > os.path.getmtime(lib) for d in deps):
        subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-x", "c++", src, "-o", lib], check=True)
    return ctypes.CDLL(lib)


def vk_blob(vk):
    out = struct.pack("<II", vk["nPublic"], vk["power"]) + b.to_lem(vk["k1"]) + b.to_lem(vk["k2"]) + b.to_lem(vk["w"])
    for nm in ("Qm", "Ql", "Qr", "Qo", "Qc", "S1", "S2", "S3"):
        out += b.g1_to_lem(vk[nm])
    return out + bytes(vk["X_2"])


def test_host_pairing_header_matches_the_oracle(fc):
    rng = random.Random(5)
    for _ in range(3):
        a, c = rng.randrange(1, b.R_MOD), rng.randrange(1, b.R_MOD)
        P, Q = b.g1_mul(b.G1_GEN, a), pg.g2_mul(pg.G2_GEN, c)
        out = ctypes.create_string_buffer(128)
        assert fc.fc_g2_mul(pg.g2_to_lem(pg.G2_GEN), c.to_bytes(32, "little"), out) == 1 and out.raw == pg.g2_to_lem(Q)
        o = ctypes.create_string_buffer(384)
        fc.fc_pairing(b.g1_to_lem(P), pg.g2_to_lem(Q), o, 1)
        assert o.raw == f12_lem(pg.pairing(P, Q))  # the GT value, bit for bit
    o = ctypes.create_string_buffer(384)
    fc.fc_pairing(bytes(64), pg.g2_to_lem(pg.G2_GEN), o, 1)
    assert o.raw == f12_lem(pg.F12_ONE)
    for n in (0, 1, 135, 136, 137, 272, 500):
        d = os.urandom(n)
        h = ctypes.create_string_buffer(32)
        fc.fc_keccak256_hd(d, n, h)
        assert h.raw == keccak256(d)


@pytest.mark.parametrize("name", FIXTURES)
def test_host_verify_header_matches_the_oracle(fc, name):
    fx, zkey, vk, proof, pub = _fixture(name)
    for label, pr, pu, expected in _mutations(proof, pub, random.Random(2)):
        pubs = b"".join(x.to_bytes(32, "little") for x in pu)
        assert fc.fc_plonk_verify(vk_blob(vk), pr, pubs or b"\0", len(pu)) == int(expected), label
    assert fc.fc_plonk_verify(vk_blob(vk), proof, bytes(32 * (len(pub) + 1)), len(pub) + 1) == 0  # nPublic mismatch


def test_calldata_text_needs_no_gpu():
    """`snarkjs zkey export soliditycalldata`: 0x<proof hex>,["0x<pub 64 hex>",..]"""
    from nzcb_circom_b200.snarkjs import plonk, proof_obj_to_bytes, proof_struct_to_obj
    from nzcb_circom_b200._lib import Proof

    fx, zkey, vk, proof, pub = _fixture("small")
    text = plonk.exportSolidityCallData(proof, pub)
    assert text == "0x" + proof.hex() + ",[" + ",".join('"0x%064x"' % p for p in pub) + "]"
    assert proof_obj_to_bytes(proof_struct_to_obj(Proof.from_buffer_copy(proof))) == proof
    assert plonk.exportSolidityCallData(json.loads(fx["proof_json"]), pub) == text


# ---------------------------------------------------------------- GPU
@pytest.mark.gpu
def test_gpu_pairing_eq_matches_the_oracle(ctx):
    from nzcb_circom_b200.ffjavascript import pairingEq
    from nzcb_circom_b200.snarkjs import powersoftau

    rng = random.Random(11)
    a, c = rng.randrange(1, b.R_MOD), rng.randrange(1, b.R_MOD)
    P, Q = b.g1_mul(b.G1_GEN, a), pg.g2_mul(pg.G2_GEN, c)
    ok, gt = pairingEq(b.g1_to_lem(P), pg.g2_to_lem(Q), ctx=ctx, want_gt=True)
    assert not ok and gt == f12_lem(pg.pairing(P, Q))                       # e(P, Q) itself, bit for bit
    aP, aQ = b.g1_mul(b.G1_GEN, a), pg.g2_mul(pg.G2_GEN, a)
    assert pairingEq(b.g1_to_lem(aP), pg.g2_to_lem(pg.G2_GEN), b.g1_to_lem(b.g1_neg(b.G1_GEN)), pg.g2_to_lem(aQ), ctx=ctx)
    assert not pairingEq(b.g1_to_lem(aP), pg.g2_to_lem(pg.G2_GEN), b.g1_to_lem(b.G1_GEN), pg.g2_to_lem(aQ), ctx=ctx)
    assert pairingEq(bytes(64), pg.g2_to_lem(Q), b.g1_to_lem(P), bytes(128), ctx=ctx)   # infinities pair to 1
    assert pairingEq(ctx=ctx)
    assert powersoftau.new_g2(TAU, ctx) == X2_LEM
    assert powersoftau.new_g2(1, ctx) == pg.g2_to_lem(pg.G2_GEN)
    from nzcb_circom_b200 import NzcbError
    with pytest.raises(NzcbError):
        pairingEq(b.g1_to_lem((1, 3)), pg.g2_to_lem(Q), ctx=ctx)  # not on the curve


@pytest.mark.gpu
@pytest.mark.parametrize("name", FIXTURES)
def test_gpu_verify_matches_the_oracle(ctx, name):
    from nzcb_circom_b200.snarkjs import VKey, plonk, zKey

    fx, zkey, vk, proof, pub = _fixture(name)
    vkj = zKey.exportVerificationKey(zkey)
    x2 = pg.g2_mul(pg.G2_GEN, TAU)
    vkj["X_2"] = [[str(x2[0][0]), str(x2[0][1])], [str(x2[1][0]), str(x2[1][1])], ["1", "0"]]
    vkey = VKey(vkj, ctx)                                   # through the verification_key.json text
    muts = _mutations(proof, pub, random.Random(2))
    got = plonk.verify_batch(vkey, [m[2] for m in muts], [m[1] for m in muts], ctx)
    assert got == [m[3] for m in muts], [m[0] for m in muts]
    for env, val in (("NZCB_VERIFY_GROUP", "16"), ("NZCB_VERIFY_GROUP", "8"), ("NZCB_VERIFY_SERIAL", "1")):
        os.environ[env] = val                                # the packed and the one-thread forms agree
        try:
            assert plonk.verify_batch(vkey, [m[2] for m in muts], [m[1] for m in muts], ctx) == got, (env, val)
        finally:
            del os.environ[env]
    assert got == [op.verify(vk, m[2], op.proof_from_bytes(m[1])) for m in muts]
    assert plonk.verify(vkey, pub, json.loads(fx["proof_json"]), ctx) is True      # the snarkjs call shape
    assert plonk.verify_batch(vkey, [pub + [0]], [proof], ctx) == [False]         # nPublic mismatch
    vkj_wrong = dict(vkj)
    x3 = pg.g2_mul(pg.G2_GEN, TAU + 1)
    vkj_wrong["X_2"] = [[str(x3[0][0]), str(x3[0][1])], [str(x3[1][0]), str(x3[1][1])], ["1", "0"]]
    assert plonk.verify(vkj_wrong, pub, proof, ctx) is False
    vkey.close()


def _twist_point_outside_g2(rng):
    """a point of the sextic twist y^2 = x^3 + 3/(9+u) that is NOT of order r (the twist's cofactor is ~2^254)"""
    q = b.P_MOD
    one, mone = (1, 0), (q - 1, 0)
    while True:
        x = (rng.randrange(q), rng.randrange(q))
        a = pg.f2_add(pg.f2_mul(pg.f2_sqr(x), x), pg.TWIST_B)
        a1 = pg.f2_pow(a, (q - 3) // 4)                      # square root in Fq2, q = 3 mod 4
        alpha = pg.f2_mul(pg.f2_sqr(a1), a)
        if pg.f2_mul(pg.f2_conj(alpha), alpha) == mone:
            continue                                         # not a square
        x0 = pg.f2_mul(a1, a)
        y = pg.f2_mul((0, 1), x0) if alpha == mone else pg.f2_mul(pg.f2_pow(pg.f2_add(one, alpha), (q - 1) // 2), x0)
        Q = (x, y)
        assert pg.g2_is_on_curve(Q)
        if pg.g2_add(pg.g2_mul(Q, b.R_MOD - 1), Q) is not None:  # [r] Q != infinity
            return Q


@pytest.mark.gpu
def test_gpu_vkey_refuses_a_key_without_a_valid_x2(ctx):
    """with X_2 at infinity e(., X_2) = 1, the pairing check degenerates to B1 == infinity and a forged
    proof (Z = Wxiw = infinity, eval_zw = 0, Wxi = -(F - E)/xi) passes.  Such a key -- and one whose X_2 lies on the
    twist outside the r-torsion -- must not load as a verification key at all."""
    from nzcb_circom_b200 import NzcbError
    from nzcb_circom_b200.snarkjs import VKey, plonk, zKey

    fx, zkey, vk, proof, pub = _fixture("small")
    with pytest.raises(NzcbError):
        VKey(zkey, ctx)                                      # the fixture zkey was written without X_2 (all zero)
    vkj = zKey.exportVerificationKey(zkey)
    vkj["X_2"] = [["0", "0"], ["0", "0"], ["0", "0"]]
    with pytest.raises(NzcbError):
        plonk.verify(vkj, pub, proof, ctx)
    Q = _twist_point_outside_g2(random.Random(3))
    vkj["X_2"] = [[str(Q[0][0]), str(Q[0][1])], [str(Q[1][0]), str(Q[1][1])], ["1", "0"]]
    with pytest.raises(NzcbError):
        plonk.verify(vkj, pub, proof, ctx)
    vkj["X_2"] = [["1", "2"], ["3", "4"], ["1", "0"]]     # not on the twist
    with pytest.raises(NzcbError):
        plonk.verify(vkj, pub, proof, ctx)


@pytest.mark.gpu
def test_gpu_nzcp_live_proofs_verify_on_the_device(nzcp_live_prover):
    """full size: prove on the GPU, verify on the GPU with the key taken from the zkey itself (X_2 = [tau]_2 from
    nzcb_srs_g2), cross-checked by the oracle's pairing verifier; a large batch with every other proof tampered"""
    from nzcb_circom_b200 import nzcp_helpers as H
    from nzcb_circom_b200.snarkjs import VKey, plonk

    pr = nzcp_live_prover
    passes = [H.synth_pass(s) for s in (30, 31)]
    res = pr.prove_passes([(p["toBeSigned"], p["data"]) for p in passes])
    assert all(s == 0 for _, _, s in res)
    pubs = [[int(x) for x in r[1]] for r in res]
    assert pr.verify(pubs, [r[0] for r in res]) == [True, True]
    vk = op.vk_from_json(pr.vk)
    assert op.vk_x2(vk) == pg.g2_mul(pg.G2_GEN, TAU)
    assert op.verify(vk, pubs[0], op.proof_from_bytes(res[0][0]))
    vkey = VKey(pr.zkey_bytes, pr.ctx)                      # straight from the zkey file bytes
    proofs, publics, want = [], [], []
    for i in range(1024):
        p = bytearray(res[i % 2][0])
        if i % 2:
            p[576 + (i % 224)] ^= 1 << (i % 8)
        proofs.append(bytes(p))
        publics.append(pubs[i % 2])
        want.append(i % 2 == 0)
    assert plonk.verify_batch(vkey, publics, proofs, pr.ctx) == want
    for group in ("16", "4", "2"):                             # several proofs per warp: the large-batch form
        os.environ["NZCB_VERIFY_GROUP"] = group
        try:
            assert plonk.verify_batch(vkey, publics[:1000], proofs[:1000], pr.ctx) == want[:1000], group
            assert plonk.verify_batch(vkey, publics[:3], proofs[:3], pr.ctx) == want[:3], group   # a ragged last group
        finally:
            del os.environ["NZCB_VERIFY_GROUP"]
    print("verify batch of 1024: %.1f ms on the device" % pr.ctx.last_device_ms)
    vkey.close()
