"""Host mirror of circom_tester 0.0.9's ``wasm`` tester (un-vendored,
/root/reference/yarn.lock:2503) -- the boundary the reference's Mocha tests
drive: ``cir = await wasm_tester(path)``; ``witness = await
cir.calculateWitness(input, sanityCheck)`` (/root/reference/test/nzcp.js:3,42,
104..347; test/cbor.js; test/quinSelector.js).

``wasm_tester("…/circuits/getV3_test.circom")`` resolves the wrapper circuit by
its file name (the 30-odd one-line ``component main = Template(args);`` files of
/root/reference/circuits/, restated in MAINS below), builds it with the eDSL
(the circom stand-in), loads the witness program into libnzcb.so and runs it on
the GPU.  A failed ``===`` / assert rejects with "Assert Failed", like the JS."""
import ctypes
import os

from ._lib import NZCB_E_ASSERT, NzcbError, default_context
from .circom import cbor, nzcp
from .circom.builder import Circuit


def _main_quin_selector(n):
    def build(c):
        out = c.output("out")
        ins = c.input("in", n) if n else []
        index = c.input("index")
        c.assign_output(out, cbor.quin_selector(c, ins, index))
    return build


def _main_get_v(n):
    def build(c):
        v = c.output("v")
        b = c.input("bytes", n)
        pos = c.input("pos")
        c.assign_output(v, cbor.get_v(c, b, pos))
    return build


def _simple(outs, ins, fn):
    """outs: names; ins: (name, dims...) in declaration order; fn(c, *inputs) -> tuple of outputs"""
    def build(c):
        o = [c.output(n, *d) for n, *d in outs]
        i = [c.input(n, *d) for n, *d in ins]
        res = fn(c, *i)
        if not isinstance(res, tuple):
            res = (res,)
        for ow, r in zip(o, res):
            if isinstance(ow, list):
                for a, b in zip(ow, r):
                    c.assign_output(a, b)
            else:
                c.assign_output(ow, r)
    return build


def _find_cwt(n):
    return _simple([("vcPos",), ("exp",)], [("mapLen",), ("bytes", n), ("pos",)],
                   lambda c, m, b, p: nzcp.find_cwt_claims(c, b, p, m, 0, 4))


def _find_cs(n):
    return _simple([("needlePos",)], [("mapLen",), ("bytes", n), ("pos",)],
                   lambda c, m, b, p: nzcp.find_cred_subj(c, b, p, m, 2, 4))


def _read_cs(n, buf):
    def fn(c, m, b, p):
        (g, gl), (f, fl), (d, dl) = nzcp.read_cred_subj(c, b, p, m, buf)
        return g, gl, f, fl, d, dl
    return _simple([("givenName", buf), ("givenNameLen",), ("familyName", buf), ("familyNameLen",), ("dob", buf),
                    ("dobLen",)], [("mapLen",), ("bytes", n), ("pos",)], fn)


# file name (without .circom) -> builder; the main components of /root/reference/circuits/*
MAINS = {
    "getType_test": _simple([("type",)], [("v",)], lambda c, v: cbor.get_type(c, v)),
    "getX_test": _simple([("x",)], [("v",)], lambda c, v: cbor.get_x(c, v)),
    "getV3_test": _main_get_v(3), "getV4_test": _main_get_v(4), "getV5_test": _main_get_v(5),
    "decodeUint32_test": _simple([("value",)], [("v",)], lambda c, v: cbor.decode_uint23(c, v)),
    "decodeUint_test": _simple([("value",), ("nextPos",)], [("v",), ("bytes", 4), ("pos",)],
                               lambda c, v, b, p: cbor.decode_uint(c, v, b, p)),
    "readType_test": _simple([("nextPos",), ("type",), ("v",)], [("bytes", 3), ("pos",)],
                             lambda c, b, p: cbor.read_type(c, b, p)),
    "skipValueScalar_test": _simple([("nextPos",)], [("bytes", 5), ("pos",)],
                                    lambda c, b, p: cbor.skip_value_scalar(c, b, p)),
    "skipValue5_test": _simple([("nextPos",)], [("bytes", 5), ("pos",)], lambda c, b, p: cbor.skip_value(c, b, p, 4)),
    "skipValue6_test": _simple([("nextPos",)], [("bytes", 6), ("pos",)], lambda c, b, p: cbor.skip_value(c, b, p, 4)),
    "stringEquals_test": _simple([("out",)], [("bytes", 5), ("pos",), ("len",)],
                                 lambda c, b, p, l: cbor.string_equals(c, b, p, l, [97, 98, 99, 100, 101])),
    "readStringLength_test": _simple([("len",), ("nextPos",)], [("bytes", 5), ("pos",)],
                                     lambda c, b, p: cbor.read_string_length(c, b, p)),
    "readMapLength_test": _simple([("len",), ("nextPos",)], [("pos",), ("bytes", 7)],
                                  lambda c, p, b: cbor.read_map_length(c, b, p)),
    "copyString_test": _simple([("outbytes", 4), ("nextPos",), ("len",)], [("bytes", 5), ("pos",)],
                               lambda c, b, p: cbor.copy_string(c, b, p, 4)),
    "constructNullifier_test": _simple(
        [("result", 64), ("resultLen",)],
        [("givenName", 64), ("givenNameLen",), ("familyName", 64), ("familyNameLen",), ("dob", 64), ("dobLen",)],
        lambda c, g, gl, f, fl, d, dl: nzcp.construct_nullifier(c, g, gl, f, fl, d, dl)),
    "findCWTClaims_exampleTest": _find_cwt(314), "findCWTClaims_liveTest": _find_cwt(351),
    "findCredSubj_exampleTest": _find_cs(314), "findCredSubj_liveTest": _find_cs(351),
    "readCredSubj_exampleTest": _read_cs(314, 32), "readCredSubj_liveTest": _read_cs(351, 64),
    "nzcp_exampleTest": lambda c: nzcp.nzcp_pub_identity(c, 0, 314, 0, 4, 2, 4),
    "nzcp_liveTest": lambda c: nzcp.nzcp_pub_identity(c, 1, 351, 0, 4, 2, 4),
    "nzcp_example": lambda c: nzcp.nzcp_pub_identity(c, 0, 314, 0, 4, 2, 4),
    "nzcp_live": lambda c: nzcp.nzcp_pub_identity(c, 1, 351, 0, 4, 2, 4),
}
for _n in range(6):
    MAINS[f"quinSelector{_n}_test"] = _main_quin_selector(_n)

_compiled = {}
_CACHE_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_cache")
_ALIAS = {"nzcp_exampleTest": "nzcp_example", "nzcp_liveTest": "nzcp_live"}  # identical mains


def _source_tag():
    import hashlib
    h = hashlib.sha256()
    d = os.path.join(os.path.dirname(os.path.abspath(__file__)), "circom")
    for f in sorted(os.listdir(d)):
        if f.endswith(".py"):
            with open(os.path.join(d, f), "rb") as fh:
                h.update(fh.read())
    with open(os.path.abspath(__file__), "rb") as fh:
        h.update(fh.read())
    return h.hexdigest()[:12]


def compile_circuit(name, use_cache=True):
    """circom stand-in: build the named main component -> Artifact (.r1cs + witness program).
    Big circuits are cached on disk under _cache/ like circom's own build outputs."""
    import pickle
    import zlib
    key = os.path.splitext(os.path.basename(name))[0]
    key = _ALIAS.get(key, key)
    if key not in MAINS:
        raise FileNotFoundError(f"no main component known for {name}")
    variant = "-di" if os.environ.get("NZCB_WITNESS_DROP_IMPLIED", "1") != "0" else ""  # builder.Circuit.assert_zero
    ckey = key + variant
    if ckey in _compiled:
        return _compiled[ckey]
    path = os.path.join(_CACHE_DIR, f"{key}-{_source_tag()}{variant}.pkz")
    if use_cache and os.path.exists(path):
        try:
            with open(path, "rb") as fh:
                _compiled[ckey] = pickle.loads(zlib.decompress(fh.read()))
            return _compiled[ckey]
        except Exception:
            pass
    c = Circuit(key)
    MAINS[key](c)
    art = c.finalize().artifact()
    _compiled[ckey] = art
    if use_cache and art.n_witness > 20000:
        os.makedirs(_CACHE_DIR, exist_ok=True)
        tmp = path + f".tmp{os.getpid()}"
        with open(tmp, "wb") as fh:
            fh.write(zlib.compress(pickle.dumps(art, protocol=4), 1))
        os.replace(tmp, path)
        for f in os.listdir(_CACHE_DIR):  # build outputs of older builder sources: never read again
            if f.startswith(key + "-") and f.endswith(".pkz") and f"-{_source_tag()}" not in f:
                try:
                    os.remove(os.path.join(_CACHE_DIR, f))
                except OSError:
                    pass
    return _compiled[ckey]


class WasmTester:
    """What ``await wasm_tester(path)`` returns."""

    def __init__(self, compiled, ctx=None):
        self.compiled = compiled
        self._wprog = None
        self._handles = {}

    @property
    def wprog(self):
        return self.compiled.wprog_bytes()

    @property
    def r1cs(self):
        return self.compiled.r1cs_bytes()

    def _handle(self, ctx):
        key = id(ctx)
        if key not in self._handles:
            data = self.compiled.wprog_bytes(native=True)  # word-level SHA-2 steps where the circuit has them
            buf = (ctypes.c_uint8 * len(data)).from_buffer_copy(data)
            h = ctypes.c_void_p()
            ctx.check(ctx.lib.nzcb_circuit_load(ctx.h, buf, len(data), ctypes.byref(h)))
            self._handles[key] = (h, ctx)
        return self._handles[key][0]

    def calculateWitnessBatch(self, inputs, sanityCheck=True, ctx=None, want_witness=True):
        """B inputs -> (raw witness bytes B x nWitness x 32 LE | None, [status])."""
        ctx = ctx or default_context()
        h = self._handle(ctx)
        n_in = self.compiled.n_in
        flat = bytearray()
        for inp in inputs:
            vals = self.compiled.flatten_input(inp) if isinstance(inp, dict) else inp
            flat += b"".join(int(v).to_bytes(32, "little") for v in vals)
        B = len(inputs)
        ibuf = (ctypes.c_uint8 * max(1, len(flat))).from_buffer_copy(bytes(flat) or b"\0")
        out = (ctypes.c_uint8 * (B * self.compiled.n_witness * 32))() if want_witness else None
        status = (ctypes.c_int32 * B)()
        ctx.check(ctx.lib.nzcb_witness_batch(ctx.h, h, ibuf, B, out, status))
        st = [int(s) if sanityCheck else 0 for s in status]
        return (bytes(out) if want_witness else None), st

    def calculateWitness(self, input, sanityCheck=True, ctx=None):
        """-> list of ints: w[0] = 1, outputs, inputs, internals.  Raises NzcbError("Assert Failed")."""
        raw, st = self.calculateWitnessBatch([input], sanityCheck, ctx)
        if st[0] != 0:
            raise NzcbError(NZCB_E_ASSERT, "Assert Failed")
        return [int.from_bytes(raw[i:i + 32], "little") for i in range(0, len(raw), 32)]

    def close(self):
        for h, ctx in self._handles.values():
            if ctx.h:
                ctx.lib.nzcb_circuit_free(h)
        self._handles = {}


def wasm_tester(path, ctx=None):
    return WasmTester(compile_circuit(path), ctx)


wasm = wasm_tester
