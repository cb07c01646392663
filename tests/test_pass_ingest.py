"""Pass ingest: pass URI -> COSE_Sign1 -> ToBeSigned -> circuit inputs.

CPU part: the oracle (oracle/pass_ingest.py, a literal restatement of test/helpers/nzcp.js with JavaScript's
number semantics) against the reference's own vectors, and the kernel's arithmetic header (csrc/ingest.cuh)
compiled for the host against the oracle.  GPU part: nzcb_pass_ingest_batch / nzcb_plonk_fullprove_uri_batch
through the C ABI against the oracle, bit for bit."""
import ctypes
import hashlib
import os
import random
import subprocess

import pytest

from nzcb_circom_b200 import nzcp_helpers as H
from oracle import pass_ingest as pi
from tests.pass_cases import cases, synth_uri

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostcheck")


def test_oracle_example_pass_matches_the_reference_vectors():
    """EXAMPLE_PASS_URI (test/nzcp.js:71) -> ToBeSigned: 314 bytes (EXAMPLE_TOBESIGNED_MAX, test/nzcp.js:11) whose
    SHA-256 is `example2` of test/utils.js:17; claims map at 27, vc at 76 (test/nzcp.js:108)"""
    tbs = pi.to_be_signed(H.EXAMPLE_PASS_URI)
    assert len(tbs) == 314
    assert hashlib.sha256(tbs).hexdigest() == "271ce33d671a2d3b816d788135f4343e14bc66802f8cd841faac939e8c11f3ee"
    assert tbs[:12] == b"\x84\x6aSignature1" and tbs[27] == 0xA5
    assert tbs[73:76] == b"\x62vc" and tbs[76] == 0xA4 and tbs[246] == 0xA3
    # the host-side mirror used by the other tests agrees
    c = H.getCOSE(H.EXAMPLE_PASS_URI)
    assert H.encodeToBeSigned(c["bodyProtected"], c["payload"]) == tbs


def test_oracle_inputs_match_the_test_marshalling():
    """circuit_inputs == flatten({toBeSigned, toBeSignedLen, data}) of test/nzcp.js:36-41 built from the helper
    mirrors (bufferToBitArray(fitBytes(..)), evmRearrangeBytes)"""
    rng = random.Random(3)
    for seed in range(4):
        uri, p = synth_uri(seed)
        data = bytes(rng.randrange(256) for _ in range(20))
        st, fitted, n, inputs = pi.ingest(uri, 351, data)
        assert st == 0 and n == len(p["toBeSigned"]) and fitted == H.fitBytes(p["toBeSigned"], 351)
        ref = H.nzcp_input(p["toBeSigned"], 351, data)
        assert inputs == list(ref["toBeSigned"]) + [ref["toBeSignedLen"]] + list(ref["data"])


def test_oracle_js_number_semantics():
    lab = dict(cases())
    ok = lambda k: pi.ingest(lab[k], 351, bytes(20))[0]
    assert ok("array_len_8byte_fold") == 0      # x << 32 is x << 0 in JS
    assert ok("array_len_neg") == -1            # new Array(negative) throws
    assert ok("unprot_empty_array") == 0 and ok("unprot_empty_bstr") == 0   # typeof 'object', no keys
    assert ok("unprot_empty_text") == -1 and ok("unprot_map1") == -1
    assert ok("other_prefix") == 0              # substring(8): prefix unchecked
    assert ok("truncated_1") == 0 and ok("truncated_2") == -1
    assert pi.base32ToBytes("A") == b"\x00"     # ceil(5n/8)-sized Uint8Array


@pytest.fixture(scope="module")
def fc():
    src = os.path.join(HERE, "fieldcheck.cpp")
    lib = os.path.join(HERE, "libfieldcheck.so")
    deps = [src] + [os.path.join(HERE, "..", "..", "nzcb_circom_b200", "csrc", f) for f in ("fp.cuh", "g1.cuh", "keccak.h", "ingest.cuh")]
    if not os.path.exists(lib) or any(os.path.getmtime(d) > os.path.getmtime(lib) for d in deps):
        subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-x", "c++", src, "-o", lib], check=True)
    return ctypes.CDLL(lib)


def test_kernel_header_matches_the_oracle_on_the_host(fc):
    n_ok = 0
    for label, uri in cases(n_synth=40):
        for max_len in (314, 351):
            d20 = bytes((7 * k + len(label)) & 255 for k in range(20))
            st, fitted, ln, inputs = pi.ingest(uri, max_len, d20)
            ub = uri.encode("latin-1")
            tbs = ctypes.create_string_buffer(max_len)
            tl = ctypes.c_uint32()
            inp = (ctypes.c_uint32 * (8 * max_len + 161))()
            rc = fc.fc_ingest(ub + bytes(16), len(ub), d20, max_len, tbs, ctypes.byref(tl), inp)
            assert (rc, tbs.raw, tl.value) == (st, fitted, ln), label
            assert list(inp) == inputs, label
            n_ok += st == 0
    assert n_ok >= 80


def test_kernel_header_fuzz_against_the_oracle(fc):
    """random COSE-like byte strings (every length form of CBOR heads, the empty-object variants of data[1], byte
    corruption, truncation, ragged base32 tails, several maxLen): header and oracle agree on every case"""
    from tests.pass_cases import b32encode

    rng = random.Random(2024)
    heads = [b"\xd2\x84", b"\xd2\x98\x04", b"\xd2\x99\x00\x04", b"\xd2\x9a\x00\x00\x00\x04", b"\xd2\x9b" + bytes(7) + b"\x04",
             b"\xd2\x83", b"\xd1\x84", b"\xd2\xa4", b"\xd2\x9f", b"\xd2\x44"]
    unprots = [b"\xa0", b"\x80", b"\x40", b"\x60", b"\xb8\x00", b"\x98\x00", b"\x58\x00", b"\xa1\x01\x02", b"\xf6", b"\x00",
               b"\x20", b"\xba\x00\x00\x00\x00", b"\xbb" + bytes(8), b"\x5b" + bytes(8), b"\xc0\xa0"]

    def bstr_any(x):
        n, k = len(x), rng.randrange(5)
        if k == 0 and n <= 23:
            return bytes([0x40 + n]) + x
        if k <= 1 and n < 256:
            return bytes([0x58, n]) + x
        if k <= 2:
            return bytes([0x59, n >> 8, n & 255]) + x
        if k == 3:
            return bytes([0x5A]) + n.to_bytes(4, "big") + x
        return bytes([0x5B]) + n.to_bytes(8, "big") + x

    accepted = 0
    for it in range(1500):
        prot = bytes(rng.randrange(256) for _ in range(rng.choice([0, 1, 10, 13, 23, 24, 30])))
        pay = bytes(rng.randrange(256) for _ in range(rng.choice([0, 5, 200, 255, 256, 300, 340, 700])))
        sig = bytes(rng.randrange(256) for _ in range(rng.choice([0, 64, 10])))
        raw = rng.choice(heads) + bstr_any(prot) + rng.choice(unprots) + bstr_any(pay) + bstr_any(sig)
        r = rng.random()
        if r < 0.3:
            raw = bytearray(raw)
            for _ in range(rng.randrange(1, 4)):
                raw[rng.randrange(len(raw))] = rng.randrange(256)
            raw = bytes(raw)
        elif r < 0.4:
            raw = raw[:rng.randrange(len(raw))]
        uri = rng.choice(["NZCP:/1/", "12345678", "NZCP:/2/"]) + b32encode(raw)
        if rng.random() < 0.1:
            uri = uri[:-1]
        max_len = rng.choice([314, 351, 64, 1024])
        d20 = bytes(rng.randrange(256) for _ in range(20))
        st, fitted, ln, inputs = pi.ingest(uri, max_len, d20)
        ub = uri.encode("latin-1")
        tbs = ctypes.create_string_buffer(max_len)
        tl = ctypes.c_uint32()
        inp = (ctypes.c_uint32 * (8 * max_len + 161))()
        rc = fc.fc_ingest(ub + bytes(16), len(ub), d20, max_len, tbs, ctypes.byref(tl), inp)
        assert (rc, tbs.raw, tl.value) == (st, fitted, ln) and list(inp) == inputs, (it, raw[:24].hex())
        accepted += st == 0
    assert 200 < accepted < 1300


# ---------------------------------------------------------------- GPU
def _check_batch(ctx, uris, max_len, datas):
    from nzcb_circom_b200.pass_ingest import toBeSignedBatch

    res, inputs = toBeSignedBatch(uris, max_len, datas, ctx=ctx, want_inputs=True)
    n_in = 8 * max_len + 161
    for i, uri in enumerate(uris):
        st, fitted, ln, vals = pi.ingest(uri, max_len, datas[i])
        assert res[i] == (st, fitted, ln), i
        blk = inputs[i * n_in * 32:(i + 1) * n_in * 32]
        assert blk == b"".join(v.to_bytes(32, "little") for v in vals), i


@pytest.mark.gpu
@pytest.mark.parametrize("max_len", [314, 351])
def test_gpu_ingest_matches_the_oracle(ctx, max_len):
    rng = random.Random(max_len)
    uris = [u for _, u in cases(n_synth=60)]
    datas = [bytes(rng.randrange(256) for _ in range(20)) for _ in uris]
    _check_batch(ctx, uris, max_len, datas)
    # the reference's example pass
    from nzcb_circom_b200.pass_ingest import toBeSigned

    fitted, n = toBeSigned(H.EXAMPLE_PASS_URI, max_len, ctx=ctx)
    assert n == 314 and hashlib.sha256(fitted[:n]).hexdigest() == "271ce33d671a2d3b816d788135f4343e14bc66802f8cd841faac939e8c11f3ee"


@pytest.mark.gpu
def test_gpu_ingest_large_batch_and_fuzz(ctx):
    """4,096 passes (more CTAs than the grid holds at once: the grid-stride path) with random corruptions"""
    rng = random.Random(99)
    base = [synth_uri(s)[0] for s in range(64)]
    uris = []
    for i in range(4096):
        u = base[i % 64]
        r = rng.random()
        if r < 0.15:    # flip one character to another base32 symbol: usually still decodes, sometimes breaks CBOR
            k = rng.randrange(8, len(u))
            u = u[:k] + rng.choice("ABCDEFGHIJKLMNOPQRSTUVWXYZ234567") + u[k + 1:]
        elif r < 0.20:
            u = u[:rng.randrange(0, len(u))]
        elif r < 0.23:
            k = rng.randrange(8, len(u))
            u = u[:k] + rng.choice("01890=_ a") + u[k + 1:]
        uris.append(u)
    datas = [bytes(rng.randrange(256) for _ in range(20)) for _ in uris]
    _check_batch(ctx, uris, 351, datas)
    from nzcb_circom_b200.pass_ingest import toBeSignedBatch

    assert toBeSignedBatch([], 351, ctx=ctx) == []


@pytest.mark.gpu
def test_gpu_fullprove_from_uris(nzcp_live_prover):
    """pass URIs -> proofs in one call: same bytes as fullProve over host-marshalled inputs; an undecodable pass
    and a pass the circuit rejects are reported apart and do not poison the batch"""
    from nzcb_circom_b200.pass_ingest import fullProveURIs
    from oracle import bn254 as b

    pr = nzcp_live_prover
    rng = random.Random(17)
    (u0, p0), (u1, p1) = synth_uri(20), synth_uri(21)
    data = [p0["data"], bytes(20), p1["data"], bytes(20)]
    broken = u1[:-40]
    # a pass longer than the circuit takes decodes fine and is rejected by the circuit (toBeSignedLen < 352)
    from tests.pass_cases import b32encode, cose_bytes
    long_uri = "NZCP:/1/" + b32encode(cose_bytes(b"\xa2\x04\x45key-1\x01\x26", bytes(400)))
    uris = [u0, broken, u1, long_uri]
    blinders = [[rng.randrange(b.R_MOD) for _ in range(9)] for _ in uris]
    res = fullProveURIs(uris, 351, pr.tester, pr.zk, data, blinders, ctx=pr.ctx)
    assert [s for _, _, s in res] == [0, -1, 0, -6]
    ref = pr.prove_passes([(p0["toBeSigned"], p0["data"]), (p1["toBeSigned"], p1["data"])], [blinders[0], blinders[2]])
    assert res[0][0] == ref[0][0] and res[2][0] == ref[1][0] and res[0][1] == ref[0][1] and res[2][1] == ref[1][1]
    nh, th, exp, d = H.nzcp_decode_outputs([int(x) for x in res[0][1]])
    assert th == hashlib.sha256(p0["toBeSigned"]).digest() and exp == p0["exp"] and d == p0["data"]
