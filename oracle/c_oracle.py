"""ctypes binding of the C restatement (oracle/c/nzcb_oracle.c).  Test infrastructure only."""
import ctypes
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "c", "liboracle.so")
_lib = None


def build():
    subprocess.run(["make", "-C", HERE, "-s"], check=True)
    return LIB


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        L = ctypes.CDLL(LIB)
        vp, sz = ctypes.c_void_p, ctypes.c_size_t
        L.oracle_ntt.argtypes = [vp, ctypes.c_uint, ctypes.c_int]
        L.oracle_msm.argtypes = [vp, vp, sz, vp]
        L.oracle_keccak256.argtypes = [vp, sz, vp]
        L.oracle_witness.argtypes = [vp, sz, vp, vp]
        L.oracle_prove.argtypes = [vp, sz, vp, sz, vp, vp, vp]
        L.oracle_fullprove.argtypes = [vp, sz, vp, vp, sz, vp, vp, vp]
        L.oracle_num_threads.restype = ctypes.c_int
        _lib = L
    return _lib


def _buf(b):
    if isinstance(b, bytes):
        return ctypes.cast(ctypes.c_char_p(b), ctypes.c_void_p)
    if isinstance(b, bytearray):
        return (ctypes.c_uint8 * len(b)).from_buffer(b)
    return b


def num_threads():
    return lib().oracle_num_threads()


def use_all_cores():
    """OpenMP threads = the cores this process may run on (torchrun sets OMP_NUM_THREADS=1 for its ranks)"""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    lib().oracle_set_num_threads(n)
    return num_threads()


def ntt(data_lem: bytes, inverse=False):
    n = len(data_lem) // 32
    buf = ctypes.create_string_buffer(data_lem, len(data_lem))
    lib().oracle_ntt(buf, n.bit_length() - 1, 1 if inverse else 0)
    return buf.raw


def msm(bases_lem: bytes, scalars_le: bytes):
    out = ctypes.create_string_buffer(64)
    lib().oracle_msm(_buf(bases_lem), _buf(scalars_le), len(scalars_le) // 32, out)
    return out.raw


def keccak256(data: bytes):
    out = ctypes.create_string_buffer(32)
    lib().oracle_keccak256(_buf(data), len(data), out)
    return out.raw


def witness(wprog: bytes, inputs_le: bytes, n_total: int):
    out = ctypes.create_string_buffer(n_total * 32)
    rc = lib().oracle_witness(_buf(wprog), len(wprog), _buf(inputs_le), out)
    return rc, out.raw


def prove(zkey, wtns: bytes, blinders, n_public: int):
    bl = b"".join(int(x).to_bytes(32, "little") for x in blinders)
    proof = ctypes.create_string_buffer(800)
    pub = ctypes.create_string_buffer(32 * max(1, n_public))
    rc = lib().oracle_prove(_buf(zkey), len(zkey), _buf(wtns), len(wtns), _buf(bl), proof, pub)
    return rc, proof.raw, [int.from_bytes(pub.raw[i * 32:(i + 1) * 32], "little") for i in range(n_public)]


def fullprove(wprog: bytes, inputs_le: bytes, zkey, blinders, n_public: int):
    bl = b"".join(int(x).to_bytes(32, "little") for x in blinders)
    proof = ctypes.create_string_buffer(800)
    pub = ctypes.create_string_buffer(32 * max(1, n_public))
    rc = lib().oracle_fullprove(_buf(wprog), len(wprog), _buf(inputs_le), _buf(zkey), len(zkey), _buf(bl), proof, pub)
    return rc, proof.raw, [int.from_bytes(pub.raw[i * 32:(i + 1) * 32], "little") for i in range(n_public)]
