"""The host-only entries on the drop-in boundary (csrc/formats.cu, include/nzcb.h): circom_runtime's input contract
(named signals, error codes 1 / 2 / 3 / 6 and "Not all inputs have been set", SURVEY.md A.4 -- what every
`cir.calculateWitness({...})` of test/nzcp.js:42, test/cbor.js, test/quinSelector.js passes through), the .wtns
writer, `zkey export verificationkey` as JSON text, the circom stand-in's command line, and a C program that links
libnzcb.so without Python.  None of these needs a GPU; the end-to-end use of the named inputs is GPU-marked."""
import base64
import ctypes
import json
import os
import struct
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
R = 21888242871839275222246405745257275088548364400416034343698204186575808495617


def _lib():
    from nzcb_circom_b200 import _lib as L
    return L.load()


def _resolve(sym, signals):
    """signals: [(name, [values])] in the order the JS object would enumerate them -> (rc, message, inputs bytes)"""
    from nzcb_circom_b200.circom.builder import fnv1a64

    lib = _lib()
    n_in = struct.unpack_from("<I", sym, 12)[0]
    hashes = (ctypes.c_uint64 * max(1, len(signals)))(*[fnv1a64(n) for n, _ in signals])
    counts = (ctypes.c_uint32 * max(1, len(signals)))(*[len(v) for _, v in signals])
    vals = b"".join((x % (1 << 256)).to_bytes(32, "little") for _, v in signals for x in v) or b"\0"
    out = ctypes.create_string_buffer(max(32, n_in * 32))
    err = ctypes.create_string_buffer(128)
    rc = lib.nzcb_inputs_resolve(sym, len(sym), len(signals), hashes, counts, vals, out, err, 128)
    return rc, err.value.decode(), out.raw[:n_in * 32]


def test_fnv1a64_matches_the_builder_and_known_answers():
    from nzcb_circom_b200.circom.builder import fnv1a64

    lib = _lib()
    assert fnv1a64("") == 0xCBF29CE484222325 and fnv1a64("a") == 0xAF63DC4C8601EC8C      # FNV-1a 64 test vectors
    for name in ("", "a", "toBeSigned", "toBeSignedLen", "data", "bytes", "pos"):
        assert lib.nzcb_fnv1a64(name.encode(), len(name)) == fnv1a64(name)


def test_named_inputs_follow_circom_runtime():
    from nzcb_circom_b200.circom_tester import compile_circuit

    art = compile_circuit("skipValue5_test")          # inputs: bytes[5], pos
    sym = art.sym_bytes()
    good = [("bytes", [0x83, 23, 23, 23, 0]), ("pos", [0])]
    rc, msg, buf = _resolve(sym, good)
    assert rc == 0 and buf == b"".join(int(v).to_bytes(32, "little") for v in art.flatten_input({"bytes": good[0][1], "pos": 0}))
    rc, _, buf2 = _resolve(sym, list(reversed(good)))  # key order of the JS object does not matter
    assert rc == 0 and buf2 == buf
    rc, _, buf3 = _resolve(sym, [("bytes", [0x83 + R, 23, 23, 23, 0]), ("pos", [R])])   # values are taken mod r (Fr.e)
    assert rc == 0 and buf3 == buf
    assert _resolve(sym, [("bytes", good[0][1]), ("pos", [0]), ("nope", [1])])[:2] == (-102, "Too many signals set")
    assert _resolve(sym, [("nope", [1]), ("pos", [0])])[:2] == (-101, "Signal not found")
    assert _resolve(sym, [("nope", [])] + good)[:2] == (-101, "Signal not found")
    assert _resolve(sym, [("bytes", [1, 2, 3, 4, 5, 6]), ("pos", [0])])[:2] == (-106, "Input signal array access exceeds the size")
    assert _resolve(sym, [("pos", [0]), ("pos", [0])])[:2] == (-103, "Signal already set")
    assert _resolve(sym, [("bytes", [1, 2, 3])])[:2] == (-107, "Not all inputs have been set. Only 3 out of 6")
    assert _resolve(sym, [])[:2] == (-107, "Not all inputs have been set. Only 0 out of 6")
    assert _resolve(b"NZSX" + sym[4:], good)[0] == -1 and _resolve(sym[:-1], good)[0] == -1


def test_wtns_export_is_the_file_the_oracle_writes():
    from oracle.binfile import read_wtns, write_wtns

    lib = _lib()
    w = [1, 5, R - 1, 0, 123456789 << 200]
    raw = b"".join(x.to_bytes(32, "little") for x in w)
    n = ctypes.c_size_t(0)
    assert lib.nzcb_wtns_export(raw, len(w), None, ctypes.byref(n)) == 0
    out = ctypes.create_string_buffer(n.value)
    assert lib.nzcb_wtns_export(raw, len(w), out, ctypes.byref(n)) == 0
    assert out.raw == bytes(write_wtns(w)) and read_wtns(out.raw) == (R, w)
    small = ctypes.c_size_t(10)
    assert lib.nzcb_wtns_export(raw, len(w), out, ctypes.byref(small)) == -1 and small.value == n.value


@pytest.mark.parametrize("name", ["tiny", "small", "nopublic"])
def test_vkey_to_json_equals_the_python_export(name):
    from nzcb_circom_b200.snarkjs import zKey

    lib = _lib()
    with open(os.path.join(ROOT, "tests", "golden", f"plonk_{name}.json")) as fh:
        zkey = base64.b64decode(json.load(fh)["zkey_b64"])
    n = ctypes.c_size_t(0)
    assert lib.nzcb_vkey_to_json(zkey, len(zkey), None, ctypes.byref(n)) == 0
    buf = ctypes.create_string_buffer(n.value)
    assert lib.nzcb_vkey_to_json(zkey, len(zkey), buf, ctypes.byref(n)) == 0
    text = buf.value.decode()
    want = zKey.exportVerificationKey(zkey)
    assert json.loads(text) == want and list(json.loads(text)) == list(want)      # same values, same key order
    assert text == json.dumps(want, indent=1)                                     # JSON.stringify(vk, null, 1) layout
    assert lib.nzcb_vkey_to_json(b"nope" + zkey[4:], len(zkey), None, ctypes.byref(n)) == -1


def test_circom_cli_writes_what_the_js_drop_in_loads(tmp_path):
    from nzcb_circom_b200.circom_tester import compile_circuit
    from oracle.binfile import read_r1cs

    out = tmp_path / "build"
    subprocess.run([sys.executable, "-m", "nzcb_circom_b200.circom", "circuits/quinSelector3_test.circom", "getV3_test", "-o", str(out)],
                   check=True, cwd=ROOT)
    for key in ("quinSelector3_test", "getV3_test"):
        art = compile_circuit(key)
        assert (out / f"{key}.wprog").read_bytes() == art.wprog_bytes(native=True)
        assert (out / f"{key}.sym").read_bytes() == art.sym_bytes()
        r = read_r1cs((out / f"{key}.r1cs").read_bytes())
        meta = json.loads((out / f"{key}.json").read_text())
        assert meta["nInputs"] == art.n_in and meta["nWitness"] == art.n_witness and len(r.constraints) == meta["nConstraints"]
        assert [i["name"] for i in meta["inputs"]] == [n for n, _, _ in art.inputs]
    with open(os.path.join(ROOT, "integration", "nzcb.js")) as f:
        js = f.read()
    assert '".wprog"' in js and '".sym"' in js and "resolveInputs" in js   # wasm(path) resolves these files by base name


def test_c_program_links_the_library_without_python(tmp_path):
    """a maintainer's first contact with the C ABI: gcc, -lnzcb, no Python anywhere"""
    src = os.path.join(ROOT, "tests", "hostcheck", "abi_smoke.c")
    exe = str(tmp_path / "abi_smoke")
    libdir = os.path.join(ROOT, "nzcb_circom_b200")
    subprocess.run(["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), src, "-o", exe,
                    "-L", libdir, "-lnzcb", f"-Wl,-rpath,{libdir}"], check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "abi_smoke ok" in r.stdout


@pytest.mark.gpu
def test_named_inputs_end_to_end_on_the_gpu(ctx):
    """calculateWitness({bytes, pos}) through nzcb_inputs_resolve + nzcb_witness_batch + nzcb_wtns_export"""
    from nzcb_circom_b200.circom_tester import wasm_tester
    from oracle.binfile import read_wtns

    cir = wasm_tester("skipValue5_test", ctx)
    art = cir.compiled
    rc, msg, buf = _resolve(art.sym_bytes(), [("pos", [0]), ("bytes", [0x83, 23, 23, 23, 0])])
    assert rc == 0, msg
    status = (ctypes.c_int32 * 1)()
    out = ctypes.create_string_buffer(art.n_witness * 32)
    ctx.check(ctx.lib.nzcb_witness_batch(ctx.h, cir._handle(ctx), buf, 1, out, status))
    assert status[0] == 0
    w = cir.calculateWitness({"bytes": [0x83, 23, 23, 23, 0], "pos": 0}, True, ctx)
    n = ctypes.c_size_t(0)
    ctx.lib.nzcb_wtns_export(out.raw, art.n_witness, None, ctypes.byref(n))
    f = ctypes.create_string_buffer(n.value)
    assert ctx.lib.nzcb_wtns_export(out.raw, art.n_witness, f, ctypes.byref(n)) == 0
    assert read_wtns(f.raw)[1] == w and w[1] == 4      # test/cbor.js: skipValue over [23, 23, 23] ends at 4
