"""Pass ingest on the GPU: the device counterpart of the reference's test helpers `getCOSE` + `encodeToBeSigned`
(/root/reference/test/helpers/nzcp.js:141-206) and of the input object test/nzcp.js:36-41 builds, for batches of
pass URIs.  (nzcp_helpers.py keeps the host-side mirror of the same helpers, as the reference's tests use them.)"""
import ctypes

from ._lib import NzcbError, Proof, as_cbuf, default_context

NZCB_E_INVALID = -1


def _pack(uris):
    raw = [u.encode("latin-1", "replace") if isinstance(u, str) else bytes(u) for u in uris]
    off = [0]
    for r in raw:
        off.append(off[-1] + len(r))
    return b"".join(raw), (ctypes.c_uint32 * len(off))(*off)


def _data_buf(data, B):
    if data is None:
        return None
    flat = b"".join(bytes(d) for d in data)
    if len(flat) != 20 * B:
        raise ValueError("data must be 20 pass-through bytes per pass")
    return as_cbuf(flat)


def toBeSignedBatch(passURIs, maxLen, data=None, ctx=None, want_inputs=False):
    """B pass URIs -> [(status, fitBytes(ToBeSigned, maxLen), true length)] and, with want_inputs, the marshalled
    main inputs (bytes, B x (8 maxLen + 161) x 32 LE).  status -1 = "invalid data" (the JS helpers throw)."""
    ctx = ctx or default_context()
    B = len(passURIs)
    blob, off = _pack(passURIs)
    n_in = 8 * maxLen + 161
    tbs = (ctypes.c_uint8 * max(1, B * maxLen))()
    lens = (ctypes.c_uint32 * max(1, B))()
    status = (ctypes.c_int32 * max(1, B))()
    inputs = (ctypes.c_uint8 * max(1, B * n_in * 32))() if want_inputs else None
    ctx.check(ctx.lib.nzcb_pass_ingest_batch(ctx.h, as_cbuf(blob or b"\0"), off, B, _data_buf(data, B), maxLen, tbs, lens,
                                             inputs, status))
    raw = bytes(tbs)
    res = [(int(status[i]), raw[i * maxLen:(i + 1) * maxLen], int(lens[i])) for i in range(B)]
    return (res, bytes(inputs)[:B * n_in * 32]) if want_inputs else res


def toBeSigned(passURI, maxLen, ctx=None):
    """encodeToBeSigned(getCOSE(passURI)) of one pass, zero-fitted to maxLen; raises like the JS helper"""
    st, fitted, n = toBeSignedBatch([passURI], maxLen, ctx=ctx)[0]
    if st != 0:
        raise NzcbError(st, "invalid data")
    return fitted, n


def fullProveURIs(passURIs, maxLen, circuit, zkey, data=None, blinders_list=None, ctx=None):
    """snarkjs.plonk.fullProve for B pass URIs, ingest included on the device:
    -> [(proof bytes | None, publicSignals, status)]"""
    ctx = ctx or zkey.ctx
    B = len(passURIs)
    blob, off = _pack(passURIs)
    bl = None
    if blinders_list is not None:
        rawb = b"".join(int(x).to_bytes(32, "little") for bs in blinders_list for x in bs)
        bl = as_cbuf(rawb)
    out = (Proof * max(1, B))()
    npub = max(1, zkey.n_public)
    pub = (ctypes.c_uint8 * (32 * npub * max(1, B)))()
    status = (ctypes.c_int32 * max(1, B))()
    ctx.check(ctx.lib.nzcb_plonk_fullprove_uri_batch(ctx.h, circuit._handle(ctx), zkey.h, as_cbuf(blob or b"\0"), off, B,
                                                     _data_buf(data, B), maxLen, bl, out, pub, status))
    res = []
    for i in range(B):
        pb = bytes(pub[i * 32 * zkey.n_public:(i + 1) * 32 * zkey.n_public])
        public = [str(int.from_bytes(pb[k * 32:(k + 1) * 32], "little")) for k in range(zkey.n_public)]
        res.append((bytes(out[i]) if status[i] == 0 else None, public, int(status[i])))
    return res
