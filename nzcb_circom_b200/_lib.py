"""ctypes binding of libnzcb.so (include/nzcb.h).  Fails loudly when the CUDA
library is missing or no B200 is usable -- there is no CPU fallback."""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NZCB_LIB_PATH") or os.path.join(HERE, "libnzcb.so")  # override: A/B builds of the library

NZCB_OK = 0
NZCB_E_INVALID = -1
NZCB_E_CUDA = -2
NZCB_E_WITNESS = -3
NZCB_E_COPY = -4
NZCB_E_DIVIDE = -5
NZCB_E_ASSERT = -6
NZCB_E_NOMEM = -7


class NzcbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"nzcb error {code}: {msg}")
        self.code = code
        self.message = msg


class Proof(ctypes.Structure):
    _fields_ = [(n, ctypes.c_uint8 * 64) for n in ("A", "B", "C", "Z", "T1", "T2", "T3", "Wxi", "Wxiw")] + [
        (n, ctypes.c_uint8 * 32) for n in ("eval_a", "eval_b", "eval_c", "eval_s1", "eval_s2", "eval_zw", "eval_r")
    ]


_lib = None

_vp = ctypes.c_void_p
_cp = ctypes.c_char_p
_sz = ctypes.c_size_t
_i32 = ctypes.c_int32
_u32 = ctypes.c_uint32

# name -> (restype, argtypes); every symbol include/nzcb.h declares
SIGNATURES = {
    "nzcb_ctx_create": (_i32, [_i32, ctypes.POINTER(_vp)]),
    "nzcb_ctx_free": (None, [_vp]),
    "nzcb_last_error": (_cp, [_vp]),
    "nzcb_launch_count": (ctypes.c_uint64, [_vp]),
    "nzcb_last_device_ms": (ctypes.c_float, [_vp]),
    "nzcb_ctx_set_msm_split": (_i32, [_vp, _i32, _i32, _vp, _vp]),
    "nzcb_microbench": (_i32, [_vp, _i32, _u32, _u32, ctypes.POINTER(ctypes.c_double)]),
    "nzcb_microbench_madd": (_i32, [_vp, _i32, _u32, _u32, ctypes.POINTER(ctypes.c_double)]),
    "nzcb_microbench_level": (_i32, [_vp, _i32, _u32, ctypes.POINTER(ctypes.c_double)]),
    "nzcb_selftest_mul": (_i32, [_vp, _u32, ctypes.POINTER(ctypes.c_uint64)]),
    "nzcb_ntt_fr": (_i32, [_vp, _vp, _u32, _i32]),
    "nzcb_msm_g1": (_i32, [_vp, _vp, _vp, _sz, _vp]),
    "nzcb_dev_alloc": (_i32, [_vp, _sz, ctypes.POINTER(_vp)]),
    "nzcb_dev_free": (_i32, [_vp, _vp]),
    "nzcb_dev_upload": (_i32, [_vp, _vp, _vp, _sz]),
    "nzcb_dev_download": (_i32, [_vp, _vp, _vp, _sz]),
    "nzcb_ntt_fr_dev": (_i32, [_vp, _vp, _u32, _i32]),
    "nzcb_msm_g1_dev": (_i32, [_vp, _vp, _vp, _sz, _vp]),
    "nzcb_srs_g1": (_i32, [_vp, _vp, _sz, _vp]),
    "nzcb_plonk_setup": (_i32, [_vp, _vp, _sz, _vp, _sz, _vp, _vp, ctypes.POINTER(_sz)]),
    "nzcb_plonk_setup_ptau": (_i32, [_vp, _vp, _sz, _vp, _sz, _vp, ctypes.POINTER(_sz)]),
    "nzcb_ptau_info": (_i32, [_vp, _sz, ctypes.POINTER(_u32), ctypes.POINTER(_u32), ctypes.POINTER(ctypes.c_uint64),
                              ctypes.POINTER(_i32)]),
    "nzcb_plonk_setup_info": (_i32, [_vp, _vp, _sz] + [ctypes.POINTER(_u32)] * 4),
    "nzcb_zkey_load": (_i32, [_vp, _vp, _sz, ctypes.POINTER(_vp)]),
    "nzcb_zkey_free": (None, [_vp]),
    "nzcb_zkey_info": (_i32, [_vp] + [ctypes.POINTER(_u32)] * 5),
    "nzcb_plonk_prove": (_i32, [_vp, _vp, _vp, _sz, _vp, ctypes.POINTER(Proof), _vp]),
    "nzcb_plonk_prove_batch": (_i32, [_vp, _vp, _vp, _vp, _vp, _sz, _vp, _vp, _vp]),
    "nzcb_proof_to_json": (_i32, [ctypes.POINTER(Proof), _vp, ctypes.POINTER(_sz)]),
    "nzcb_circuit_load": (_i32, [_vp, _vp, _sz, ctypes.POINTER(_vp)]),
    "nzcb_circuit_free": (None, [_vp]),
    "nzcb_circuit_info": (_i32, [_vp] + [ctypes.POINTER(_u32)] * 3),
    "nzcb_witness_batch": (_i32, [_vp, _vp, _vp, _sz, _vp, _vp]),
    "nzcb_nccl_unique_id": (_i32, [_cp, _vp]),
    "nzcb_ctx_set_msm_split_nccl": (_i32, [_vp, _i32, _i32, _cp, _vp]),
    "nzcb_fnv1a64": (ctypes.c_uint64, [_cp, _sz]),
    "nzcb_inputs_resolve": (_i32, [_vp, _sz, _u32, _vp, _vp, _vp, _vp, _vp, _sz]),
    "nzcb_wtns_export": (_i32, [_vp, _u32, _vp, ctypes.POINTER(_sz)]),
    "nzcb_vkey_to_json": (_i32, [_vp, _sz, _vp, ctypes.POINTER(_sz)]),
    "nzcb_witness_batch_ex": (_i32, [_vp, _vp, _vp, _sz, _vp, _vp, _sz, _vp, _vp]),
    "nzcb_witness_batch_ex_dev": (_i32, [_vp, _vp, _vp, _sz, _vp, _vp, _sz, _vp, _vp]),
    "nzcb_plonk_fullprove_batch": (_i32, [_vp, _vp, _vp, _vp, _sz, _vp, _vp, _vp, _vp]),
    "nzcb_plonk_fullprove_batch_dev": (_i32, [_vp, _vp, _vp, _vp, _sz, _vp, _vp, _vp, _vp]),
    "nzcb_vkey_from_zkey": (_i32, [_vp, _vp, _sz, ctypes.POINTER(_vp)]),
    "nzcb_vkey_from_json": (_i32, [_vp, _cp, _sz, ctypes.POINTER(_vp)]),
    "nzcb_vkey_free": (None, [_vp]),
    "nzcb_plonk_verify_batch": (_i32, [_vp, _vp, _vp, _vp, _u32, _sz, _vp]),
    "nzcb_pairing_eq": (_i32, [_vp, _vp, _vp, _u32, ctypes.POINTER(_i32), _vp]),
    "nzcb_srs_g2": (_i32, [_vp, _vp, _vp]),
    "nzcb_proof_to_calldata": (_i32, [ctypes.POINTER(Proof), _vp, _u32, _vp, ctypes.POINTER(_sz)]),
    "nzcb_pass_ingest_batch": (_i32, [_vp, _vp, _vp, _sz, _vp, _u32, _vp, _vp, _vp, _vp]),
    "nzcb_pass_ingest_batch_dev": (_i32, [_vp, _vp, _vp, _sz, _vp, _u32, _vp, _vp]),
    "nzcb_plonk_fullprove_uri_batch": (_i32, [_vp, _vp, _vp, _vp, _vp, _sz, _vp, _u32, _vp, _vp, _vp, _vp]),
    "nzcb_g1_table_create": (_i32, [_vp, _vp, _sz, ctypes.POINTER(_vp)]),
    "nzcb_g1_table_free": (None, [_vp]),
    "nzcb_msm_g1_table": (_i32, [_vp, _vp, _vp, _sz, _vp]),
    "nzcb_msm_g1_table_dev": (_i32, [_vp, _vp, _vp, _vp, _i32, _vp]),
    "nzcb_g1_lagrange_basis": (_i32, [_vp, _vp, _u32, _vp]),
    "nzcb_profile": (_i32, [_vp, _i32]),
    "nzcb_profile_entries": (_i32, [_vp, ctypes.POINTER(ctypes.c_double)]),
    "nzcb_profile_read": (_i32, [_vp, ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(ctypes.c_double),
                                 ctypes.POINTER(ctypes.c_double)]),
}


def load():
    """dlopen libnzcb.so and bind every declared symbol (no GPU needed for this)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NzcbError(NZCB_E_CUDA, f"{LIB_PATH} is missing: run `python __graft_entry__.py build` "
                                     "(nvcc, sm_100a).  There is no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError = header/library mismatch: fail loudly
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class Context:
    """One GPU, one stream set (nzcb_ctx)."""

    def __init__(self, device_id=0):
        self.lib = load()
        h = _vp()
        rc = self.lib.nzcb_ctx_create(device_id, ctypes.byref(h))
        if rc != 0:
            raise NzcbError(rc, self.lib.nzcb_last_error(None).decode())
        self.h = h

    def check(self, rc):
        if rc != 0:
            raise NzcbError(rc, self.lib.nzcb_last_error(self.h).decode())

    @property
    def launches(self):
        return int(self.lib.nzcb_launch_count(self.h))

    @property
    def last_device_ms(self):
        return float(self.lib.nzcb_last_device_ms(self.h))

    def close(self):
        if getattr(self, "h", None):
            self.lib.nzcb_ctx_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_msm_split_nccl(self, rank, world, dist=None):
        """Latency mode with the exchange on the device (nzcb_ctx_set_msm_split_nccl): the library opens its own NCCL
        communicator; `dist` (torch.distributed, any backend) only carries the 128-byte unique id from rank 0.
        world <= 1 switches the mode off."""
        if world <= 1:
            self.check(self.lib.nzcb_ctx_set_msm_split_nccl(self.h, 0, 1, None, None))
            return
        path = nccl_library_path()
        ids = [None]
        if rank == 0:  # a failure here is broadcast too, so that no rank is left waiting in the collective
            buf = (ctypes.c_uint8 * 128)()
            if self.lib.nzcb_nccl_unique_id(path, buf) == 0:
                ids[0] = bytes(buf)
        dist.broadcast_object_list(ids, src=0)
        if ids[0] is None:
            raise NzcbError(NZCB_E_INVALID, f"cannot load NCCL ({path!r}) for the device-side exchange")
        idbuf = (ctypes.c_uint8 * 128).from_buffer_copy(ids[0])
        self.check(self.lib.nzcb_ctx_set_msm_split_nccl(self.h, rank, world, path, idbuf))

    ALLGATHER_FN = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t)

    def set_msm_split(self, rank, world, allgather=None):
        """Latency mode: this ctx commits slice `rank` of `world` of every fixed-base MSM.  allgather(send: bytes) ->
        bytes of all ranks in rank order (world x len(send)); None / world 1 switches the mode off."""
        if world <= 1 or allgather is None:
            self._split_cb = None
            self.check(self.lib.nzcb_ctx_set_msm_split(self.h, 0, 1, None, None))
            return

        def _cb(_user, send, recv, nbytes):
            try:
                out = allgather(ctypes.string_at(send, nbytes))
                if len(out) != world * nbytes:
                    return 1
                ctypes.memmove(recv, out, len(out))
                return 0
            except Exception:  # never unwind through the C frames
                import traceback
                traceback.print_exc()
                return 1

        self._split_cb = Context.ALLGATHER_FN(_cb)  # keep the thunk alive as long as the mode is on
        self.check(self.lib.nzcb_ctx_set_msm_split(self.h, rank, world, ctypes.cast(self._split_cb, ctypes.c_void_p), None))

    def microbench(self, kind, iters=2000, blocks_per_sm=8):
        v = ctypes.c_double()
        self.check(self.lib.nzcb_microbench(self.h, kind, iters, blocks_per_sm, ctypes.byref(v)))
        return v.value

    def microbench_madd(self, variant, iters=2000, log_table=10):
        v = ctypes.c_double()
        self.check(self.lib.nzcb_microbench_madd(self.h, variant, iters, log_table, ctypes.byref(v)))
        return v.value

    def microbench_level(self, kind, iters=20000):
        v = ctypes.c_double()
        self.check(self.lib.nzcb_microbench_level(self.h, kind, iters, ctypes.byref(v)))
        return v.value

    def profile(self, enable=True):
        self.check(self.lib.nzcb_profile(self.h, 1 if enable else 0))

    def profile_entries(self):
        """bucket additions actually executed by the timed launches (read before profile_read)"""
        v = ctypes.c_double()
        self.check(self.lib.nzcb_profile_entries(self.h, ctypes.byref(v)))
        return v.value

    def profile_read(self):
        """(launches, total device ms, algorithmic modmul) of the MSM bucket-accumulation kernel"""
        n, ms, mm = ctypes.c_uint64(), ctypes.c_double(), ctypes.c_double()
        self.check(self.lib.nzcb_profile_read(self.h, ctypes.byref(n), ctypes.byref(ms), ctypes.byref(mm)))
        return n.value, ms.value, mm.value

    def selftest_mul(self, n=1 << 20):
        v = ctypes.c_uint64()
        self.check(self.lib.nzcb_selftest_mul(self.h, n, ctypes.byref(v)))
        return v.value

    # --- device buffers (bench / roofline path) ---
    def dev_alloc(self, nbytes):
        p = _vp()
        self.check(self.lib.nzcb_dev_alloc(self.h, nbytes, ctypes.byref(p)))
        return p

    def dev_free(self, p):
        self.check(self.lib.nzcb_dev_free(self.h, p))

    def dev_upload(self, p, data):
        buf = (ctypes.c_uint8 * len(data)).from_buffer_copy(data) if not isinstance(data, ctypes.Array) else data
        self.check(self.lib.nzcb_dev_upload(self.h, p, buf, len(data)))

    def dev_download(self, p, nbytes):
        buf = (ctypes.c_uint8 * nbytes)()
        self.check(self.lib.nzcb_dev_download(self.h, buf, p, nbytes))
        return bytes(buf)


def nccl_library_path():
    """the libnccl.so.2 that torch bundles (nvidia-nccl wheel), as bytes for the C ABI; None = loader search path"""
    try:
        import nvidia.nccl as _n
        for base in list(getattr(_n, "__path__", [])) + [os.path.dirname(getattr(_n, "__file__", None) or "")]:
            p = os.path.join(base, "lib", "libnccl.so.2")
            if base and os.path.exists(p):
                return p.encode()
        return None
    except Exception:
        return None


def as_cbuf(data):
    """zero-copy read-only view of a bytes object for the C ABI (borrowed for the call only)"""
    if isinstance(data, ctypes.Array):
        return data
    if isinstance(data, bytes):
        return ctypes.cast(ctypes.c_char_p(data), ctypes.c_void_p)
    if isinstance(data, bytearray):
        return (ctypes.c_uint8 * len(data)).from_buffer(data)
    return (ctypes.c_uint8 * len(data)).from_buffer_copy(bytes(data))


_default_ctx = {}


def default_context(device_id=None):
    if device_id is None:
        device_id = int(os.environ.get("LOCAL_RANK", "0"))
    if device_id not in _default_ctx:
        _default_ctx[device_id] = Context(device_id)
    return _default_ctx[device_id]
