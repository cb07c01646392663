"""CBOR-walk templates of the reference, restated on the builder eDSL.
Each function cites the template it follows in /root/reference/circuits/.
Signals that the circom source defines by ``<==`` from a *linear* expression are
carried as LCs (what ``--O2`` leaves after substitution); every quadratic ``<==``
and every hint is one wire, in source order."""
from .builder import LC, OP_QUINSEL, Circuit
from .circomlib import (bits2num, calculate_total, is_equal, is_zero, less_than, log2, num2bits, shr)

MAJOR_TYPE_INT = 0
MAJOR_TYPE_STRING = 3
MAJOR_TYPE_ARRAY = 4
MAJOR_TYPE_MAP = 5


def quin_selector(c: Circuit, ins, index):
    """QuinSelector(choices) -- circuits/quinSelector.circom:11-42"""
    choices = len(ins)
    index = LC.of(index)
    if choices > 0:
        # a multi-term index would be re-expanded into PLONK additions by every one of the
        # 2*choices constraints below; bind it to one signal first
        if len(index.t) > 1:
            index = c.wire(index)
        bits = log2(choices) + 1
        lt = less_than(c, bits, index, choices)          # :19-24
        c.assert_eq(lt, 1)
    p0, w0 = len(c.prog), c.n_wires
    chain = c.chain()                                     # sums[i] <== eqs[i] * in[i] + sums[i-1]
    eqs, sums = [], []
    for i in range(choices):
        eq = is_zero(c, i - index)                        # :32-33
        eqs.append(eq)
        sums.append(chain.step(eq, ins[i]))               # :37
    out = chain.finish()                                  # :41  (0 when choices == 0)
    # The whole selector is also ONE word-level instruction of the native witness program: per choice the IsZero
    # inverse, the equality flag and -- where the input is not a constant, so that the running sum is a signal --
    # the sum.  Only when every signal allocated above is one of those (a signal index, nothing folded away).
    if choices > 0 and not index.is_const():
        eq_w = [e.single_wire() for e in eqs]
        sum_w, seen = [], set()
        for sacc in sums:
            w = sacc.single_wire()
            fresh = w is not None and w >= w0 and w not in seen
            if fresh:
                seen.add(w)
            sum_w.append(w if fresh else None)
        if all(w is not None and w > w0 for w in eq_w) and \
                c.n_wires - w0 == 2 * choices + sum(1 for w in sum_w if w is not None) and \
                len(set(eq_w) | {w - 1 for w in eq_w} | seen) == c.n_wires - w0:
            c.fuse(p0, OP_QUINSEL, {"w0": w0, "index": index, "ins": [LC.of(x) for x in ins], "eq_w": eq_w, "sum_w": sum_w})
    return out


def get_type(c, v):
    """GetType -- cbortpl.circom:26-52: v >> 5"""
    bits = num2bits(c, v, 8)
    return bits2num(shr(bits, 5)[:3])


def get_x(c, v):
    """GetX -- cbortpl.circom:57-73: v & 31"""
    bits = num2bits(c, v, 8)
    return bits2num(bits[:5])


def get_v(c, bytes_, pos):
    """GetV(BytesLen) -- cbortpl.circom:79-90"""
    return quin_selector(c, bytes_, pos)


def decode_uint23(c, v):
    """DecodeUint23 -- cbortpl.circom:95-109"""
    x = get_x(c, v)
    lt = less_than(c, 8, x, 24)
    c.assert_eq(lt, 1)
    return x


def decode_uint(c, v, bytes_, pos):
    """DecodeUint(BytesLen) -- cbortpl.circom:115-237.  Returns (value, nextPos)."""
    pos = LC.of(pos)
    x = get_x(c, v)
    cond23 = less_than(c, 8, x, 24)
    cond24 = is_equal(c, x, 24)
    cond25 = is_equal(c, x, 25)
    cond26 = is_equal(c, x, 26)
    value23, next23 = x, pos
    value24 = get_v(c, bytes_, c.mul(cond24, pos))                     # :162
    next24 = pos + 1
    v1_25 = get_v(c, bytes_, c.mul(cond25, pos)) * 256                  # :174-176
    v2_25 = get_v(c, bytes_, c.mul(cond25, pos + 1))                    # :178-180
    value25 = v1_25 + v2_25
    next25 = pos + 2
    v1_26 = get_v(c, bytes_, c.mul(cond26, pos)) * 16777216             # :199-201
    v2_26 = get_v(c, bytes_, c.mul(cond26, pos + 1)) * 65536
    v3_26 = get_v(c, bytes_, c.mul(cond26, pos + 2)) * 256
    v4_26 = get_v(c, bytes_, c.mul(cond26, pos + 3))
    value26 = v1_26 + v2_26 + v3_26 + v4_26
    next26 = pos + 4
    value = calculate_total([c.mul(cond23, value23), c.mul(cond24, value24), c.mul(cond25, value25),
                             c.mul(cond26, value26)])                    # :224-229
    next_pos = calculate_total([c.mul(cond23, next23), c.mul(cond24, next24), c.mul(cond25, next25),
                                c.mul(cond26, next26)])                  # :231-236
    return value, next_pos


def read_type(c, bytes_, pos):
    """ReadType(BytesLen) -- cbortpl.circom:243-261.  Returns (nextPos, type, v)."""
    v = get_v(c, bytes_, pos)
    return LC.of(pos) + 1, get_type(c, v), v


def skip_value_scalar(c, bytes_, pos):
    """SkipValueScalar(BytesLen) -- cbortpl.circom:266-299"""
    nxt, typ, v = read_type(c, bytes_, pos)
    value, dnext = decode_uint(c, v, bytes_, nxt)
    is_int = is_equal(c, typ, MAJOR_TYPE_INT)
    is_str = is_equal(c, typ, MAJOR_TYPE_STRING)
    return calculate_total([c.mul(is_int, dnext), c.mul(is_str, dnext + value)])


def skip_value(c, bytes_, pos, max_array_len):
    """SkipValue(BytesLen, MaxArrayLen) -- cbortpl.circom:306-368"""
    nxt, typ, v = read_type(c, bytes_, pos)
    value, dnext = decode_uint(c, v, bytes_, nxt)
    is_int = is_equal(c, typ, MAJOR_TYPE_INT)
    is_str = is_equal(c, typ, MAJOR_TYPE_STRING)
    is_arr = is_equal(c, typ, MAJOR_TYPE_ARRAY)
    next_pos_array = []
    bits = log2(max_array_len) + 1
    arr_len = c.mul(is_arr, value) if max_array_len else None
    for i in range(max_array_len):
        lt = less_than(c, bits, i, arr_len)                              # :343-346
        should = c.mul(is_arr, lt)
        start = dnext if i == 0 else next_pos_array[i - 1]
        np_i = skip_value_scalar(c, bytes_, c.mul(start, should))        # :348-352
        next_pos_array.append(np_i)
    qs = quin_selector(c, next_pos_array, c.mul(is_arr, value - 1))      # :354
    return calculate_total([c.mul(is_int, dnext), c.mul(is_str, dnext + value), c.mul(is_arr, qs)])


def string_equals(c, bytes_, pos, length, const_bytes):
    """StringEquals(BytesLen, ConstBytes, ConstBytesLen) -- cbortpl.circom:375-411"""
    assert len(const_bytes) <= len(bytes_)
    pos = LC.of(pos)
    cond_sum = is_equal(c, length, len(const_bytes))
    for i, ch in enumerate(const_bytes):
        v = get_v(c, bytes_, pos + i)
        cond_sum = cond_sum + is_equal(c, ch, v)
    return is_zero(c, (len(const_bytes) + 1) - cond_sum)


def read_string_length(c, bytes_, pos):
    """ReadStringLength(BytesLen) -- cbortpl.circom:417-437.  Returns (len, nextPos)."""
    nxt, typ, v = read_type(c, bytes_, pos)
    c.assert_eq(typ, MAJOR_TYPE_STRING)                                  # hardcore_assert :429
    value, _ = decode_uint(c, v, bytes_, nxt)
    return value, nxt


def read_map_length(c, bytes_, pos):
    """ReadMapLength(BytesLen) -- cbortpl.circom:443-462.  Returns (len, nextPos)."""
    nxt, typ, v = read_type(c, bytes_, pos)
    c.assert_eq(typ, MAJOR_TYPE_MAP)                                     # hardcore_assert :455
    return decode_uint23(c, v), nxt


def copy_string(c, bytes_, pos, max_len):
    """CopyString(BytesLen, MaxLen) -- cbortpl.circom:469-504.  Returns (outbytes, nextPos, len)."""
    assert max_len <= len(bytes_)
    length, nxt = read_string_length(c, bytes_, pos)
    bits = log2(max_len) + 1
    out = []
    for i in range(max_len):
        v = get_v(c, bytes_, nxt + i)
        lt = less_than(c, bits, i, length)
        out.append(c.mul(v, lt))
    return out, nxt + length, length
