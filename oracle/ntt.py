"""Fr NTT -- oracle restatement of ffjavascript ``Fr.fft`` / ``Fr.ifft``.

SURVEY.md A.1: natural order in, evaluations at w[log N]^i natural order out;
ifft includes the 1/N factor.  Test infrastructure only.
"""
from .bn254 import R_MOD, fr_root, fr_inv


def _bitrev(i, bits):
    r = 0
    for _ in range(bits):
        r = (r << 1) | (i & 1)
        i >>= 1
    return r


def _ntt(vals, w):
    n = len(vals)
    bits = n.bit_length() - 1
    assert 1 << bits == n
    a = [0] * n
    for i in range(n):
        a[_bitrev(i, bits)] = vals[i]
    m = 1
    s = 0
    while m < n:
        wm = pow(w, n >> (s + 1), R_MOD)
        for k in range(0, n, 2 * m):
            t = 1
            for j in range(m):
                u = a[k + j]
                v = a[k + j + m] * t % R_MOD
                a[k + j] = (u + v) % R_MOD
                a[k + j + m] = (u - v) % R_MOD
                t = t * wm % R_MOD
        m *= 2
        s += 1
    return a


def fft(vals):
    n = len(vals)
    return _ntt(vals, fr_root(n.bit_length() - 1))


def ifft(vals):
    n = len(vals)
    w = fr_inv(fr_root(n.bit_length() - 1))
    ninv = fr_inv(n)
    return [x * ninv % R_MOD for x in _ntt(vals, w)]


def dft_naive(vals):
    """O(n^2) definition, to pin fft() itself on tiny sizes."""
    n = len(vals)
    w = fr_root(n.bit_length() - 1)
    return [sum(v * pow(w, i * j, R_MOD) for j, v in enumerate(vals)) % R_MOD for i in range(n)]
