"""Port of /root/reference/test/nzcp.js and test/utils.js.  The embedded example pass
(test/nzcp.js:71) is the golden vector; the live passes of the reference come from .env secrets
that are not available, so the live circuits run on synthetic passes built to the same layout
 with expectations from hashlib, like the reference takes them from Node crypto."""
import hashlib

import pytest

from nzcb_circom_b200 import nzcp_helpers as H
from tests.wbackend import BACKENDS, calc

pytestmark = pytest.mark.parametrize("backend", BACKENDS)

COSE = H.getCOSE(H.EXAMPLE_PASS_URI)
EXAMPLE_TBS = H.encodeToBeSigned(COSE["bodyProtected"], COSE["payload"])
GOLDEN_TBS_SHA256 = "271ce33d671a2d3b816d788135f4343e14bc66802f8cd841faac939e8c11f3ee"  # test/utils.js:17
GOLDEN_OUT = [8464235439336389695359576364537904521787463454426143836621154307990710930,
              334204042160295982690797293769892102755483197293558786265320143920457223185,
              430989588176824417852954207888075491695208395262355815151652761069951123456]


def test_utils_kat(backend):  # test/utils.js:6-20
    if backend != "oracle":
        pytest.skip("pure host helper")
    chunks = [366677313775235426412199931337625106565467678080892143469223808086055532772, 119]
    bits = [b for c in chunks for b in H.chunkToBits(c, 248)]
    assert H.fitBytes(H.bitArrayToBuffer(bits), 32)[:32].hex() == GOLDEN_TBS_SHA256
    assert len(EXAMPLE_TBS) == 314 and hashlib.sha256(EXAMPLE_TBS).hexdigest() == GOLDEN_TBS_SHA256


def test_find_cwt_claims_example(backend):  # :98-109 (pos 28 -> vcPos 76, exp)
    w = calc("findCWTClaims_exampleTest", {"mapLen": 5, "bytes": list(H.fitBytes(EXAMPLE_TBS, 314)), "pos": 28}, backend)
    assert (w[1], w[2]) == (76, 1951416330)


def test_find_cwt_claims_live(backend):  # :111-139 (pos 31 -> vcPos 80)
    p = H.synth_pass(1)
    w = calc("findCWTClaims_liveTest", {"mapLen": 5, "bytes": list(H.fitBytes(p["toBeSigned"], 351)), "pos": 31}, backend)
    assert (w[1], w[2]) == (80, p["exp"])


def test_find_cred_subj_example(backend):  # :155-166 (pos 77 -> 246)
    w = calc("findCredSubj_exampleTest", {"mapLen": 4, "bytes": list(H.fitBytes(EXAMPLE_TBS, 314)), "pos": 77}, backend)
    assert w[1] == 246


def test_find_cred_subj_live(backend):  # :168-199 (findCredSubj_liveTest: pos 81 -> 250) on synthetic live-layout passes
    for seed in (3, 8):
        p = H.synth_pass(seed)
        w = calc("findCredSubj_liveTest", {"mapLen": 4, "bytes": list(H.fitBytes(p["toBeSigned"], 351)), "pos": 81}, backend)
        assert w[1] == 250


def _check_cred_subj(w, buf, given, family, dob):  # testReadCredSubj :201-229
    assert w[1:1 + buf] == H.padArray(H.stringToArray(given), buf) and w[1 + buf] == len(given)
    assert w[2 + buf:2 + 2 * buf] == H.padArray(H.stringToArray(family), buf) and w[2 + 2 * buf] == len(family)
    assert w[3 + 2 * buf:3 + 3 * buf] == H.padArray(H.stringToArray(dob), buf) and w[3 + 3 * buf] == len(dob)


def test_read_cred_subj_example(backend):  # :231-242 (pos 247, buffers of 32)
    w = calc("readCredSubj_exampleTest", {"mapLen": 3, "bytes": list(H.fitBytes(EXAMPLE_TBS, 314)), "pos": 247}, backend)
    _check_cred_subj(w, 32, "Jack", "Sparrow", "1960-04-16")


def test_read_cred_subj_live(backend):  # :244-270 (pos 251, buffers of 64)
    p = H.synth_pass(2)
    g, f, d = p["nullifier"].split(",")
    w = calc("readCredSubj_liveTest", {"mapLen": 3, "bytes": list(H.fitBytes(p["toBeSigned"], 351)), "pos": 251}, backend)
    _check_cred_subj(w, 64, g, f, d)
    # mapLen != 3 is rejected (hardcore_assert nzcptpl.circom:261)
    assert calc("readCredSubj_liveTest", {"mapLen": 4, "bytes": list(H.fitBytes(p["toBeSigned"], 351)), "pos": 251},
                backend) is None


@pytest.mark.parametrize("given,family,dob", [("Jack", "Sparrow", "1960-04-16"), ("A", "B", "2000-01-01"),
                                              ("Bartholomew-Maximilian", "Featherstonehaugh", "1999-12-31")])
def test_construct_nullifier(backend, given, family, dob):  # testNullifier :272-295
    n = 64
    w = calc("constructNullifier_test", {
        "givenName": H.padArray(H.stringToArray(given), n), "givenNameLen": len(given),
        "familyName": H.padArray(H.stringToArray(family), n), "familyNameLen": len(family),
        "dob": H.padArray(H.stringToArray(dob), n), "dobLen": len(dob)}, backend)
    exp = f"{given},{family},{dob}"
    assert w[1:1 + n] == H.padArray(H.stringToArray(exp), n) and w[1 + n] == len(exp)


def _check_identity(w, tbs, nullifier, exp, data):  # testNZCPPubIdentity :33-69
    nh, th, e, d = H.nzcp_decode_outputs(w[1:4])
    assert nh == hashlib.sha512(H.fitBytes(nullifier.encode(), 64)).digest()[:32]
    assert th == hashlib.sha256(tbs).digest()
    assert e == exp and d == data


def test_nzcp_example(backend):  # :328-338 + golden outputs; the witness satisfies the whole R1CS
    data = bytes(range(1, 21))
    w = calc("nzcp_exampleTest", H.nzcp_input(EXAMPLE_TBS, 314, data), backend, check_r1cs=(backend == "oracle"))
    assert w[1:4] == GOLDEN_OUT
    _check_identity(w, EXAMPLE_TBS, "Jack,Sparrow,1960-04-16", 1951416330, data)
    # circom's observable wire order: 1, outputs, inputs in declaration order (nzcptpl.circom:486-489)
    inp = H.nzcp_input(EXAMPLE_TBS, 314, data)
    assert w[0] == 1 and w[4:4 + 2512] == inp["toBeSigned"] and w[4 + 2512] == 314 and w[4 + 2513:4 + 2513 + 160] == inp["data"]


def test_nzcp_live(backend):  # :342-368 on a synthetic live-layout pass
    p = H.synth_pass(3)
    w = calc("nzcp_liveTest", H.nzcp_input(p["toBeSigned"], 351, p["data"]), backend)
    _check_identity(w, p["toBeSigned"], p["nullifier"], p["exp"], p["data"])


def test_nzcp_live_rejects(backend):
    p = H.synth_pass(4)
    good = H.nzcp_input(p["toBeSigned"], 351, p["data"])
    bad_bit = dict(good, toBeSigned=[2] + good["toBeSigned"][1:])          # non-boolean input bit (:493-496)
    assert calc("nzcp_liveTest", bad_bit, backend) is None
    too_long = dict(good, toBeSignedLen=352)                                 # toBeSignedLen < 352 (:500-505)
    assert calc("nzcp_liveTest", too_long, backend) is None
    not_map = bytearray(p["toBeSigned"])
    not_map[30] = 0x65                                                       # claims header not a map (cbortpl :455)
    assert calc("nzcp_liveTest", H.nzcp_input(bytes(not_map), 351, p["data"]), backend) is None
