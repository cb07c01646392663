"""Host mirror of the two ffjavascript 0.2.48 primitives on the proving path
(un-vendored dependency, /root/reference/yarn.lock:3905):

    curve.Fr.fft(buff) / curve.Fr.ifft(buff)     -> Fr.fft / Fr.ifft
    curve.G1.multiExpAffine(buffBases, buffScalars) -> G1.multiExpAffine

Same buffer conventions as the JS: field elements 32-byte little-endian
Montgomery, G1 affine 64 bytes, scalars 32-byte little-endian canonical.
Everything runs in libnzcb.so on the GPU."""
import ctypes

from ._lib import default_context


class _Fr:
    n8 = 32

    def fft(self, buff, ctx=None):
        return self._ntt(buff, 0, ctx)

    def ifft(self, buff, ctx=None):
        return self._ntt(buff, 1, ctx)

    @staticmethod
    def _ntt(buff, inverse, ctx):
        ctx = ctx or default_context()
        n = len(buff) // 32
        if n * 32 != len(buff) or n & (n - 1) or n == 0:
            raise ValueError("Size must be multiple of 2")  # ffjavascript's message
        log_n = n.bit_length() - 1
        data = (ctypes.c_uint8 * len(buff)).from_buffer_copy(buff)
        ctx.check(ctx.lib.nzcb_ntt_fr(ctx.h, data, log_n, inverse))
        return bytes(data)


class _G1:
    def multiExpAffine(self, buffBases, buffScalars, ctx=None):
        ctx = ctx or default_context()
        n = len(buffBases) // 64
        if len(buffScalars) != n * 32:
            raise ValueError("Number of scalars does not match number of bases")
        out = (ctypes.c_uint8 * 64)()
        b = (ctypes.c_uint8 * max(1, len(buffBases))).from_buffer_copy(buffBases or b"\0")
        s = (ctypes.c_uint8 * max(1, len(buffScalars))).from_buffer_copy(buffScalars or b"\0")
        ctx.check(ctx.lib.nzcb_msm_g1(ctx.h, b, s, n, out))
        return bytes(out)


Fr = _Fr()
G1 = _G1()


class G1Table:
    """Fixed bases with precomputed window shifts (nzcb_g1_table_*): how the prover holds the zkey's SRS.
    multiExpAffine over the first n bases; `batch` runs up to four MSMs as one sort + accumulation."""

    def __init__(self, buffBases, ctx=None):
        self.ctx = ctx or default_context()
        self.n = len(buffBases) // 64
        h = ctypes.c_void_p()
        b = (ctypes.c_uint8 * len(buffBases)).from_buffer_copy(buffBases)
        self.ctx.check(self.ctx.lib.nzcb_g1_table_create(self.ctx.h, b, self.n, ctypes.byref(h)))
        self.h = h

    def multiExpAffine(self, buffScalars):
        n = len(buffScalars) // 32
        out = (ctypes.c_uint8 * 64)()
        s = (ctypes.c_uint8 * max(1, len(buffScalars))).from_buffer_copy(buffScalars or b"\0")
        self.ctx.check(self.ctx.lib.nzcb_msm_g1_table(self.ctx.h, self.h, s, n, out))
        return bytes(out)

    def batch(self, scalar_buffers):
        """[scalars bytes, ...] (<= 4) -> [affine LEM bytes, ...]"""
        K = len(scalar_buffers)
        ctx = self.ctx
        dptrs = []
        for sb in scalar_buffers:
            d = ctx.dev_alloc(max(32, len(sb)))
            if sb:
                ctx.dev_upload(d, sb)
            dptrs.append(d)
        arr = (ctypes.c_void_p * K)(*[d.value if isinstance(d, ctypes.c_void_p) else d for d in dptrs])
        ns = (ctypes.c_size_t * K)(*[len(sb) // 32 for sb in scalar_buffers])
        out = (ctypes.c_uint8 * (64 * K))()
        try:
            ctx.check(ctx.lib.nzcb_msm_g1_table_dev(ctx.h, self.h, arr, ns, K, out))
        finally:
            for d in dptrs:
                ctx.dev_free(d)
        return [bytes(out[64 * k:64 * (k + 1)]) for k in range(K)]

    def close(self):
        if getattr(self, "h", None) and self.ctx.h:
            self.ctx.lib.nzcb_g1_table_free(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def pairingEq(*args, ctx=None, want_gt=False):
    """curve.pairingEq(P_0, Q_0, P_1, Q_1, ...): prod e(P_i, Q_i) == 1.  P_i: 64 B G1 affine LEM, Q_i: 128 B G2 affine
    LEM (x.c0 x.c1 y.c0 y.c1).  want_gt: also return the GT product (384 B, coefficients of 1, w, .., w^5)."""
    ctx = ctx or default_context()
    if len(args) % 2:
        raise ValueError("Pairing arguments must be even")  # ffjavascript's message
    n = len(args) // 2
    g1 = b"".join(bytes(args[2 * i]) for i in range(n))
    g2 = b"".join(bytes(args[2 * i + 1]) for i in range(n))
    if len(g1) != 64 * n or len(g2) != 128 * n:
        raise ValueError("pairingEq takes 64-byte G1 and 128-byte G2 points")
    res = ctypes.c_int32(0)
    gt = (ctypes.c_uint8 * 384)()
    b1 = (ctypes.c_uint8 * max(1, len(g1))).from_buffer_copy(g1 or b"\0")
    b2 = (ctypes.c_uint8 * max(1, len(g2))).from_buffer_copy(g2 or b"\0")
    ctx.check(ctx.lib.nzcb_pairing_eq(ctx.h, b1, b2, n, ctypes.byref(res), gt))
    return (bool(res.value), bytes(gt)) if want_gt else bool(res.value)
