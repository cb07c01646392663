"""Keccak-256 (original Keccak padding 0x01, NOT NIST SHA-3) -- oracle.

Restates js-sha3 0.8.0 ``keccak256`` (/root/reference/yarn.lock:5074), which
snarkjs uses for the Fiat-Shamir transcript (SURVEY.md A.1 ``hashToFr``).
KAT: keccak256(b"") = c5d24601...5d85a470.  Test infrastructure only.
"""
from .bn254 import R_MOD

_RC = []
_ROT = [[0] * 5 for _ in range(5)]


def _init():
    # round constants via the LFSR of the Keccak spec
    r = 1
    for _ in range(24):
        rc = 0
        for j in range(7):
            if r & 1:
                rc ^= 1 << ((1 << j) - 1)
            r = ((r << 1) ^ ((r >> 7) * 0x71)) & 0xFF
        _RC.append(rc)
    x, y = 1, 0
    for t in range(24):
        _ROT[x][y] = ((t + 1) * (t + 2) // 2) % 64
        x, y = y, (2 * x + 3 * y) % 5


_init()
_M64 = (1 << 64) - 1


def _rol(v, n):
    n %= 64
    return ((v << n) | (v >> (64 - n))) & _M64 if n else v


def _f1600(A):
    for rnd in range(24):
        C = [A[x][0] ^ A[x][1] ^ A[x][2] ^ A[x][3] ^ A[x][4] for x in range(5)]
        D = [C[(x - 1) % 5] ^ _rol(C[(x + 1) % 5], 1) for x in range(5)]
        A = [[A[x][y] ^ D[x] for y in range(5)] for x in range(5)]
        B = [[0] * 5 for _ in range(5)]
        for x in range(5):
            for y in range(5):
                B[y][(2 * x + 3 * y) % 5] = _rol(A[x][y], _ROT[x][y])
        A = [[B[x][y] ^ ((~B[(x + 1) % 5][y]) & B[(x + 2) % 5][y]) for y in range(5)] for x in range(5)]
        A[0][0] ^= _RC[rnd]
    return A


def keccak256(data: bytes) -> bytes:
    rate = 136
    msg = bytearray(data)
    msg.append(0x01)
    while len(msg) % rate:
        msg.append(0)
    msg[-1] |= 0x80
    A = [[0] * 5 for _ in range(5)]
    for off in range(0, len(msg), rate):
        blk = msg[off:off + rate]
        for i in range(rate // 8):
            A[i % 5][i // 5] ^= int.from_bytes(blk[8 * i:8 * i + 8], "little")
        A = _f1600(A)
    out = b"".join(A[i % 5][i // 5].to_bytes(8, "little") for i in range(4))
    return out


def hash_to_fr(data: bytes) -> int:
    """snarkjs ``hashToFr``: big-endian digest reduced mod r."""
    return int.from_bytes(keccak256(data), "big") % R_MOD
