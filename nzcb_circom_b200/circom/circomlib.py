"""circomlib gadgets the reference circuits include (un-vendored: they sit inside the
sha256-var-circom zip fetched at build time, /root/reference/Makefile:20-25; include
sites circuits/cbortpl.circom:3-4, circuits/nzcptpl.circom:3-4,
circuits/quinSelector.circom:4).  Semantics per SURVEY.md A.6."""
from .builder import LC, Circuit


def log2(x):
    """circuits/log2.circom:5-12 -- floor(log2 x), log2(0) = -1"""
    z = -1
    while x:
        z += 1
        x //= 2
    return z


def pow_(x, y):
    """circuits/pow.circom:4-10"""
    return 1 if y == 0 else x * pow_(x, y - 1)


def num2bits(c: Circuit, x, n):
    """Num2Bits(n): out[i] <-- (in >> i) & 1; out[i]*(out[i]-1) === 0; sum out[i] 2^i === in"""
    x = LC.of(x)
    bits = c.hint_bits(x, n)
    acc = LC()
    for i, b in enumerate(bits):
        c.assert_zero(b * (b - 1), implied=True)  # b was just written by the decomposition
        acc = acc + b * (1 << i)
    c.assert_eq(acc, x)
    return bits


def bits2num(bits):
    """Bits2Num(n): sum in[i] 2^i (LSB first) -- linear"""
    acc = LC()
    for i, b in enumerate(bits):
        acc = acc + LC.of(b) * (1 << i)
    return acc


def is_zero(c: Circuit, x):
    """IsZero: inv <-- in != 0 ? 1/in : 0; out <== -in*inv + 1; in*out === 0"""
    x = LC.of(x)
    if x.is_const():
        return LC(None, 1 if x.k == 0 else 0)
    inv = c.hint_inv(x)
    out = c.quad((-x) * inv + 1)
    c.assert_zero(x * out, implied=True)  # holds for inv = 1/in or 0, which is what the hint computes
    return out


def is_equal(c: Circuit, a, b):
    """IsEqual: IsZero(in[1] - in[0])"""
    return is_zero(c, LC.of(b) - LC.of(a))


def less_than(c: Circuit, n, a, b):
    """LessThan(n): Num2Bits(n+1)(in[0] + 2^n - in[1]); out = 1 - bit n"""
    assert n <= 252
    bits = num2bits(c, LC.of(a) + (1 << n) - LC.of(b), n + 1)
    return 1 - bits[n]


def calculate_total(nums):
    """CalculateTotal(n): running sum -- linear"""
    acc = LC()
    for x in nums:
        acc = acc + x
    return acc


def shr(bits, r):
    """ShR(n, r): out[i] = in[i + r] or 0"""
    n = len(bits)
    return [bits[i + r] if i + r < n else LC() for i in range(n)]
