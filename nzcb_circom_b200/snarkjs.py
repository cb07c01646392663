"""Host mirror of the snarkjs 0.4.12 entry points on the proving path
(un-vendored npm dependency, /root/reference/yarn.lock:7279; the reference's
own use is the Makefile recipe /root/reference/Makefile:54-62 and the
``wasm_tester`` call sites of test/nzcp.js).  Same names and argument meaning
as the JS:

    plonk.setup(r1cs, ptau)            -> zkey bytes        (`snarkjs plonk setup`)
    plonk.prove(zkey, wtns)            -> (proof, publicSignals)
    plonk.fullProve(input, circuit, zkey)
    wtns.calculate(input, circuit)     -> .wtns bytes

``zkey`` / ``wtns`` are the file *bytes* ({type:"mem"} in JS) or a path;
``proof`` is the dict snarkjs returns (decimal strings).  Everything below the
C ABI runs on the GPU in libnzcb.so; errors surface as NzcbError carrying
snarkjs' message.  Deterministic proofs: pass ``blinders`` (nine ints, the
values ``Fr.random()`` would have returned in round 1, b1..b9).
"""
import ctypes
import json
import struct

from ._lib import NzcbError, Proof, as_cbuf, default_context

R_MOD = 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001

_POINTS = ("A", "B", "C", "Z", "T1", "T2", "T3", "Wxi", "Wxiw")
_EVALS = ("eval_a", "eval_b", "eval_c", "eval_s1", "eval_s2", "eval_zw", "eval_r")


def _bytes_of(x):
    if isinstance(x, bytes):
        return x
    if isinstance(x, bytearray):
        return x
    if isinstance(x, memoryview):
        return bytes(x)
    with open(x, "rb") as f:
        return f.read()


class ZKey:
    """A PLONK proving key made device resident once per process (nzcb_zkey_load)."""

    def __init__(self, zkey, ctx=None):
        self.ctx = ctx or default_context()
        data = _bytes_of(zkey)
        h = ctypes.c_void_p()
        self.ctx.check(self.ctx.lib.nzcb_zkey_load(self.ctx.h, as_cbuf(data), len(data), ctypes.byref(h)))
        self.h = h
        vals = [ctypes.c_uint32() for _ in range(5)]
        self.ctx.lib.nzcb_zkey_info(self.h, *[ctypes.byref(v) for v in vals])
        self.n_vars, self.n_public, self.domain_size, self.n_additions, self.n_constraints = (v.value for v in vals)

    def close(self):
        if getattr(self, "h", None) and self.ctx.h:
            self.ctx.lib.nzcb_zkey_free(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_zkey_cache = {}


def _as_zkey(zkey, ctx):
    if isinstance(zkey, ZKey):
        return zkey
    if isinstance(zkey, str):  # path: cache like a long-lived prover process would
        key = (zkey, id(ctx))
        if key not in _zkey_cache:
            _zkey_cache[key] = ZKey(zkey, ctx)
        return _zkey_cache[key]
    return ZKey(zkey, ctx)


def proof_struct_to_obj(ps: Proof):
    """nzcb_proof -> the object snarkjs returns from plonk.prove (decimal strings)."""
    obj = {}

    def pt(raw):
        raw = bytes(raw)
        if raw == bytes(64):
            return ["0", "1", "0"]
        return [str(int.from_bytes(raw[:32], "big")), str(int.from_bytes(raw[32:], "big")), "1"]

    for k in ("A", "B", "C", "Z", "T1", "T2", "T3"):
        obj[k] = pt(getattr(ps, k))
    for k in _EVALS:
        obj[k] = str(int.from_bytes(bytes(getattr(ps, k)), "big"))
    obj["Wxi"] = pt(ps.Wxi)
    obj["Wxiw"] = pt(ps.Wxiw)
    obj["protocol"] = "plonk"
    obj["curve"] = "bn128"
    return obj


def proof_obj_to_bytes(obj):
    """the proof object of plonk.prove / proof.json -> the 800-byte nzcb_proof layout.  Coordinates and evaluations
    must fit 32 bytes; range and curve checks are the verifier's."""
    def pt(v):
        if str(v[2]) == "0":
            return bytes(64)
        return int(v[0]).to_bytes(32, "big") + int(v[1]).to_bytes(32, "big")

    out = b"".join(pt(obj[k]) for k in ("A", "B", "C", "Z", "T1", "T2", "T3", "Wxi", "Wxiw"))
    return out + b"".join(int(obj[k]).to_bytes(32, "big") for k in _EVALS)


class VKey:
    """a verification key resident on the GPU (nzcb_vkey); built from the vk object / JSON text of
    `snarkjs zkey export verificationkey` or straight from zkey bytes"""

    def __init__(self, vk, ctx=None):
        import json as _json

        self.ctx = ctx or default_context()
        h = ctypes.c_void_p()
        if isinstance(vk, (bytes, bytearray, memoryview)) and bytes(vk[:4]) == b"zkey":
            self.ctx.check(self.ctx.lib.nzcb_vkey_from_zkey(self.ctx.h, as_cbuf(bytes(vk)), len(vk), ctypes.byref(h)))
        else:
            text = vk if isinstance(vk, str) else (_json.dumps(vk, indent=1) if isinstance(vk, dict) else bytes(vk).decode())
            raw = text.encode()
            self.ctx.check(self.ctx.lib.nzcb_vkey_from_json(self.ctx.h, raw, len(raw), ctypes.byref(h)))
        self.h = h

    def close(self):
        if self.h and self.ctx.h:
            self.ctx.lib.nzcb_vkey_free(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _blinder_buf(blinders):
    if blinders is None:
        return None
    if len(blinders) != 9:
        raise ValueError("blinders must be the nine values b1..b9")
    raw = b"".join(int(x).to_bytes(32, "little") for x in blinders)
    return (ctypes.c_uint8 * len(raw)).from_buffer_copy(raw)


class _Plonk:
    def prove(self, zkey, wtns, blinders=None, ctx=None, raw=False):
        """snarkjs.plonk.prove(zkeyFile, wtnsFile) -> (proof, publicSignals)."""
        ctx = ctx or (zkey.ctx if isinstance(zkey, ZKey) else default_context())
        zk = _as_zkey(zkey, ctx)
        w = _bytes_of(wtns)
        wbuf = (ctypes.c_uint8 * len(w)).from_buffer_copy(w)
        ps = Proof()
        pub = (ctypes.c_uint8 * (32 * max(1, zk.n_public)))()
        ctx.check(ctx.lib.nzcb_plonk_prove(ctx.h, zk.h, wbuf, len(w), _blinder_buf(blinders), ctypes.byref(ps), pub))
        public = [str(int.from_bytes(bytes(pub[i * 32:(i + 1) * 32]), "little")) for i in range(zk.n_public)]
        if raw:
            return bytes(ps), public
        return proof_struct_to_obj(ps), public

    def prove_batch(self, zkey, wtns_list, blinders_list=None, ctx=None):
        """B independent proofs on this process's GPU; returns [(proof bytes | None, publicSignals, status)]."""
        ctx = ctx or (zkey.ctx if isinstance(zkey, ZKey) else default_context())
        zk = _as_zkey(zkey, ctx)
        B = len(wtns_list)
        bufs = [(ctypes.c_uint8 * len(w)).from_buffer_copy(w) for w in wtns_list]
        ptrs = (ctypes.c_void_p * B)(*[ctypes.addressof(b) for b in bufs])
        lens = (ctypes.c_size_t * B)(*[len(w) for w in wtns_list])
        bl = None
        if blinders_list is not None:
            raw = b"".join(int(x).to_bytes(32, "little") for bs in blinders_list for x in bs)
            bl = (ctypes.c_uint8 * len(raw)).from_buffer_copy(raw)
        out = (Proof * B)()
        pub = (ctypes.c_uint8 * (32 * max(1, zk.n_public) * B))()
        status = (ctypes.c_int32 * B)()
        ctx.check(ctx.lib.nzcb_plonk_prove_batch(ctx.h, zk.h, ptrs, lens, bl, B, out, pub, status))
        res = []
        for i in range(B):
            pb = bytes(pub[i * 32 * zk.n_public:(i + 1) * 32 * zk.n_public])
            public = [str(int.from_bytes(pb[k * 32:(k + 1) * 32], "little")) for k in range(zk.n_public)]
            res.append((bytes(out[i]) if status[i] == 0 else None, public, int(status[i])))
        return res

    def verify(self, vk, publicSignals, proof, ctx=None):
        """snarkjs.plonk.verify(vk_verifier, publicSignals, proof) -> bool.  vk: a VKey, the vk object / JSON text or
        zkey bytes; proof: the proof object or the 800 raw bytes."""
        return self.verify_batch(vk, [publicSignals], [proof], ctx)[0]

    def verify_batch(self, vk, publics_list, proofs, ctx=None):
        """B proofs against one key -> [bool]; one warp per proof on the GPU"""
        own = not isinstance(vk, VKey)
        vkey = VKey(vk, ctx) if own else vk
        ctx = vkey.ctx
        try:
            B = len(proofs)
            if B == 0:
                return []
            n_pub = len(publics_list[0])
            if any(len(p) != n_pub for p in publics_list):
                raise ValueError("every proof of a batch must carry the same number of public signals")
            raw = b"".join(p if isinstance(p, (bytes, bytearray)) else proof_obj_to_bytes(p) for p in proofs)
            if len(raw) != 800 * B:
                raise ValueError("a raw proof is 800 bytes")
            pubs = b"".join((int(x) % R_MOD).to_bytes(32, "little") for ps in publics_list for x in ps)  # snarkjs: Fr.e(x)
            valid = (ctypes.c_int32 * B)()
            ctx.check(ctx.lib.nzcb_plonk_verify_batch(ctx.h, vkey.h, as_cbuf(raw), as_cbuf(pubs or b"\0"), n_pub, B, valid))
            return [bool(v) for v in valid]
        finally:
            if own:
                vkey.close()

    def exportSolidityCallData(self, proof, publicSignals):
        """`snarkjs zkey export soliditycalldata` for a plonk proof (/root/reference/Makefile:57,62 export the
        verifier this text is fed to)"""
        from ._lib import load

        lib = load()
        raw = proof if isinstance(proof, (bytes, bytearray)) else proof_obj_to_bytes(proof)
        ps = Proof.from_buffer_copy(raw)
        pubs = b"".join((int(x) % R_MOD).to_bytes(32, "little") for x in publicSignals)
        n = ctypes.c_size_t(0)
        lib.nzcb_proof_to_calldata(ctypes.byref(ps), as_cbuf(pubs or b"\0"), len(publicSignals), None, ctypes.byref(n))
        buf = ctypes.create_string_buffer(n.value)
        rc = lib.nzcb_proof_to_calldata(ctypes.byref(ps), as_cbuf(pubs or b"\0"), len(publicSignals), buf, ctypes.byref(n))
        if rc != 0:
            raise NzcbError(rc, "proof_to_calldata failed")
        return buf.value.decode()

    def proof_json(self, proof_bytes, ctx=None):
        """proof.json text exactly as snarkjs writes it (JSON.stringify(proof, null, 1))."""
        ctx = ctx or default_context()
        ps = Proof.from_buffer_copy(proof_bytes)
        n = ctypes.c_size_t(0)
        ctx.lib.nzcb_proof_to_json(ctypes.byref(ps), None, ctypes.byref(n))
        buf = ctypes.create_string_buffer(n.value)
        rc = ctx.lib.nzcb_proof_to_json(ctypes.byref(ps), buf, ctypes.byref(n))
        if rc != 0:
            raise NzcbError(rc, "proof_to_json failed")
        return buf.value.decode()

    def setup(self, r1cs, srs_g1_lem, x2_g2_lem, ctx=None):
        """`snarkjs plonk setup circuit.r1cs pot.ptau circuit.zkey` (the ptau reduced to its
        tauG1 points, affine LEM, as section 2 of the .ptau holds them, and X_2 = tauG2[1], 128 bytes LEM).
        X_2 is required: a key written with X_2 = 0 makes every verifier that trusts it accept forged proofs
        (plonk.verify here refuses such a key).  `allow_missing_x2` is the explicit way to write a prove-only key."""
        ctx = ctx or default_context()
        if x2_g2_lem is None or len(x2_g2_lem) != 128:
            raise ValueError("plonk setup: X_2 ([tau]_2, 128 bytes) is required")
        if not any(x2_g2_lem) and not getattr(self, "allow_missing_x2", False):
            raise ValueError("plonk setup: X_2 is all zero (the point at infinity); the key could not be verified against")
        r = _bytes_of(r1cs)
        rbuf, sbuf = as_cbuf(r), as_cbuf(bytes(srs_g1_lem))
        x2 = (ctypes.c_uint8 * 128).from_buffer_copy(x2_g2_lem)
        n = ctypes.c_size_t(0)
        ctx.check(ctx.lib.nzcb_plonk_setup(ctx.h, rbuf, len(r), sbuf, len(srs_g1_lem) // 64, x2, None, ctypes.byref(n)))
        out = bytearray(n.value)
        obuf = (ctypes.c_uint8 * n.value).from_buffer(out)
        ctx.check(ctx.lib.nzcb_plonk_setup(ctx.h, rbuf, len(r), sbuf, len(srs_g1_lem) // 64, x2, obuf, ctypes.byref(n)))
        del obuf
        return out

    def setup_ptau(self, r1cs, ptau, ctx=None):
        """`snarkjs plonk setup circuit.r1cs pot.ptau circuit.zkey` with the .ptau file (bytes / path)"""
        ctx = ctx or default_context()
        r, p = _bytes_of(r1cs), _bytes_of(ptau)
        rbuf, pbuf = as_cbuf(r), as_cbuf(p)
        n = ctypes.c_size_t(0)
        ctx.check(ctx.lib.nzcb_plonk_setup_ptau(ctx.h, rbuf, len(r), pbuf, len(p), None, ctypes.byref(n)))
        out = bytearray(n.value)
        obuf = (ctypes.c_uint8 * n.value).from_buffer(out)
        ctx.check(ctx.lib.nzcb_plonk_setup_ptau(ctx.h, rbuf, len(r), pbuf, len(p), obuf, ctypes.byref(n)))
        del obuf
        return out

    def setup_info(self, r1cs, ctx=None):
        """(nGates, nAdditions, plonkNVars, power) of the R1CS -> PLONK expansion"""
        ctx = ctx or default_context()
        r = _bytes_of(r1cs)
        v = [ctypes.c_uint32() for _ in range(4)]
        ctx.check(ctx.lib.nzcb_plonk_setup_info(ctx.h, as_cbuf(r), len(r), *[ctypes.byref(x) for x in v]))
        return tuple(x.value for x in v)

    def fullProve(self, input, circuit, zkey, blinders=None, ctx=None, raw=False):
        """snarkjs.plonk.fullProve(input, wasmFile, zkeyFile): witness program and prover fused on
        the GPU (the wires never leave HBM).  `circuit` is a circom_tester.WasmTester."""
        res = self.fullProveBatch([input], circuit, zkey, None if blinders is None else [blinders], ctx)
        proof, public, status = res[0]
        if status != 0:
            c = ctx or default_context()
            raise NzcbError(status, "Assert Failed" if status == -6 else c.lib.nzcb_last_error(c.h).decode())
        return (proof, public) if raw else (proof_struct_to_obj(Proof.from_buffer_copy(proof)), public)

    def fullProveBatch(self, inputs, circuit, zkey, blinders_list=None, ctx=None):
        """B passes -> [(proof bytes | None, publicSignals, status)]; a rejected pass never fails the batch."""
        ctx = ctx or (zkey.ctx if isinstance(zkey, ZKey) else default_context())
        zk = _as_zkey(zkey, ctx)
        art = circuit.compiled
        h = circuit._handle(ctx)
        flat = bytearray()
        for inp in inputs:
            vals = art.flatten_input(inp) if isinstance(inp, dict) else inp
            flat += b"".join(int(v).to_bytes(32, "little") for v in vals)
        return self.fullProveRaw(bytes(flat), len(inputs), h, zk, blinders_list, ctx)

    def fullProveRaw(self, inputs_le, B, circuit_handle, zk, blinders_list=None, ctx=None, device_inputs=None):
        """inputs already marshalled: B x nInputs x 32 B canonical LE (host bytes), or
        device_inputs = a device pointer holding the same bytes (nzcb_dev_upload)"""
        ctx = ctx or zk.ctx
        ibuf = as_cbuf(inputs_le or b"\0") if device_inputs is None else device_inputs
        fn = ctx.lib.nzcb_plonk_fullprove_batch if device_inputs is None else ctx.lib.nzcb_plonk_fullprove_batch_dev
        bl = None
        if blinders_list is not None:
            raw = b"".join(int(x).to_bytes(32, "little") for bs in blinders_list for x in bs)
            bl = (ctypes.c_uint8 * len(raw)).from_buffer_copy(raw)
        out = (Proof * B)()
        npub = max(1, zk.n_public)
        pub = (ctypes.c_uint8 * (32 * npub * B))()
        status = (ctypes.c_int32 * B)()
        ctx.check(fn(ctx.h, circuit_handle, zk.h, ibuf, B, bl, out, pub, status))
        res = []
        for i in range(B):
            pb = bytes(pub[i * 32 * zk.n_public:(i + 1) * 32 * zk.n_public])
            public = [str(int.from_bytes(pb[k * 32:(k + 1) * 32], "little")) for k in range(zk.n_public)]
            res.append((bytes(out[i]) if status[i] == 0 else None, public, int(status[i])))
        return res


P_MOD = 0x30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47


class _ZKeyTools:
    def exportVerificationKey(self, zkey):
        """`snarkjs zkey export verificationkey` (/root/reference/Makefile:56,61): the vk object
        snarkjs writes to verification_key.json, from the zkey header (section 2) alone."""
        data = _bytes_of(zkey) if not isinstance(zkey, (bytes, bytearray)) else zkey
        if data[:4] != b"zkey":
            raise ValueError("zkey file: bad magic")
        nsec = struct.unpack_from("<I", data, 8)[0]
        pos = 12
        hdr = None
        for _ in range(nsec):
            sid, size = struct.unpack_from("<IQ", data, pos)
            pos += 12
            if sid == 2:
                hdr = data[pos:pos + size]
                break
            pos += size
        if hdr is None:
            raise ValueError("zkey file: no header section")
        rinv_r = pow(1 << 256, -1, R_MOD)
        rinv_q = pow(1 << 256, -1, P_MOD)

        def fr(b):
            return int.from_bytes(b, "little") * rinv_r % R_MOD

        def fq(b):
            return int.from_bytes(b, "little") * rinv_q % P_MOD

        def g1(b):
            if b == bytes(64):
                return ["0", "1", "0"]
            return [str(fq(b[:32])), str(fq(b[32:])), "1"]

        n_vars, n_public, domain, n_add, n_cons = struct.unpack_from("<IIIII", hdr, 72)
        power = domain.bit_length() - 1
        vk = {"protocol": "plonk", "curve": "bn128", "nPublic": n_public, "power": power, "k1": str(fr(hdr[92:124])),
              "k2": str(fr(hdr[124:156]))}
        off = 156
        for nm in ("Qm", "Ql", "Qr", "Qo", "Qc", "S1", "S2", "S3"):
            vk[nm] = g1(hdr[off:off + 64])
            off += 64
        x2 = hdr[off:off + 128]
        vk["X_2"] = [[str(fq(x2[0:32])), str(fq(x2[32:64]))], [str(fq(x2[64:96])), str(fq(x2[96:128]))], ["1", "0"]]
        w = pow(5, (R_MOD - 1) >> 28, R_MOD)
        for _ in range(28 - power):
            w = w * w % R_MOD
        vk["w"] = str(w)
        return vk


class _Powersoftau:
    def new_g1(self, tau, count, ctx=None):
        """[tau^i]G1 for i < count as affine LEM bytes: the tauG1 section of an insecure,
        known-trapdoor `snarkjs powersoftau new` (Makefile:64-67 role; synthetic inputs only)."""
        ctx = ctx or default_context()
        t = (ctypes.c_uint8 * 32).from_buffer_copy(int(tau % R_MOD).to_bytes(32, "little"))
        out = (ctypes.c_uint8 * (64 * count))()
        ctx.check(ctx.lib.nzcb_srs_g1(ctx.h, t, count, out))
        return bytes(out)


    def info(self, ptau):
        """header of a .ptau: {power, ceremonyPower, nTauG1, prepared}; raises ValueError on a malformed file"""
        from ._lib import load

        p = _bytes_of(ptau)
        v = [ctypes.c_uint32(), ctypes.c_uint32(), ctypes.c_uint64(), ctypes.c_int32()]
        if load().nzcb_ptau_info(as_cbuf(p), len(p), *[ctypes.byref(x) for x in v]) != 0:
            raise ValueError("not a bn128 .ptau file")
        return {"power": v[0].value, "ceremonyPower": v[1].value, "nTauG1": v[2].value, "prepared": bool(v[3].value)}

    def new_g2(self, tau, ctx=None):
        """[tau]G2 as 128 affine LEM bytes: the tauG2 point of the same insecure SRS, X_2 of the zkey header"""
        ctx = ctx or default_context()
        t = (ctypes.c_uint8 * 32).from_buffer_copy(int(tau % R_MOD).to_bytes(32, "little"))
        out = (ctypes.c_uint8 * 128)()
        ctx.check(ctx.lib.nzcb_srs_g2(ctx.h, t, out))
        return bytes(out)

    def lagrange_g1(self, srs_g1_lem, power, ctx=None):
        """[L_i(tau)]G1 for the 2^power domain from the tauG1 points: the Lagrange section `snarkjs powersoftau
        prepare phase2` appends to a .ptau (sections 12-15)."""
        ctx = ctx or default_context()
        n = 1 << power
        out = (ctypes.c_uint8 * (64 * n))()
        ctx.check(ctx.lib.nzcb_g1_lagrange_basis(ctx.h, as_cbuf(bytes(srs_g1_lem[:64 * n])), power, out))
        return bytes(out)


class _Wtns:
    def calculate(self, input, circuit, ctx=None):
        """snarkjs.wtns.calculate(input, wasmFile, {type:"mem"}) -> .wtns bytes.  `circuit` is a
        compiled circuit (nzcb_circom_b200.circom_tester.WasmTester)."""
        w = circuit.calculateWitness(input, True, ctx=ctx)
        return write_wtns(w)


def write_wtns(witness):
    """.wtns v2: header section 1, values section 2, canonical LE."""
    hdr = struct.pack("<I", 32) + R_MOD.to_bytes(32, "little") + struct.pack("<I", len(witness))
    body = b"".join(int(x).to_bytes(32, "little") for x in witness)
    out = b"wtns" + struct.pack("<II", 2, 2)
    out += struct.pack("<IQ", 1, len(hdr)) + hdr
    out += struct.pack("<IQ", 2, len(body)) + body
    return out


def wtns_from_raw(raw_le: bytes):
    """wrap nWitness x 32 B canonical LE values (nzcb_witness_batch output) as a .wtns file"""
    n = len(raw_le) // 32
    hdr = struct.pack("<I", 32) + R_MOD.to_bytes(32, "little") + struct.pack("<I", n)
    out = b"wtns" + struct.pack("<II", 2, 2)
    out += struct.pack("<IQ", 1, len(hdr)) + hdr
    out += struct.pack("<IQ", 2, len(raw_le)) + raw_le
    return out


plonk = _Plonk()
zKey = _ZKeyTools()
powersoftau = _Powersoftau()
wtns = _Wtns()
__all__ = ["plonk", "powersoftau", "wtns", "zKey", "ZKey", "write_wtns", "wtns_from_raw", "json"]
