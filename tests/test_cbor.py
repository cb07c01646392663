"""Port of /root/reference/test/cbor.js (every describe / it), run against the CPU oracle
and (marked gpu) the CUDA witness interpreter.  Line references are to that file."""
import pytest

from nzcb_circom_b200.nzcp_helpers import encodeArray, encodeInt, encodeMap, encodeString, padArray, stringToArray
from tests.wbackend import BACKENDS, calc, calc_many

pytestmark = pytest.mark.parametrize("backend", BACKENDS)


def test_get_type(backend):  # :10-22  exhaustive
    ws = calc_many("getType_test", [{"v": v} for v in range(255, -1, -1)], backend, check_r1cs=True)
    assert [w[1] for w in ws] == [v >> 5 for v in range(255, -1, -1)]


def test_get_x(backend):  # :24-36
    ws = calc_many("getX_test", [{"v": v} for v in range(255, -1, -1)], backend, check_r1cs=True)
    assert [w[1] for w in ws] == [v & 31 for v in range(255, -1, -1)]


@pytest.mark.parametrize("n", [3, 4, 5])
def test_get_v(backend, n):  # :38-104
    b = list(range(1, n + 1))
    ws = calc_many(f"getV{n}_test", [{"bytes": b, "pos": p} for p in range(n)], backend, check_r1cs=True)
    assert [w[1] for w in ws] == b


def test_decode_uint23(backend):  # :107-127: x <= 23 decodes, anything else is rejected
    ws = calc_many("decodeUint32_test", [{"v": v} for v in range(256)], backend)
    for v, w in enumerate(ws):
        x = v & 31
        if x <= 23:
            assert w[1] == x
        else:
            assert w is None


@pytest.mark.parametrize("bytes_,v,exp", [  # :129-178
    ([0, 0, 0, 0], 167, 7), ([0, 0, 0, 0], 168, 8),
    ([31, 0, 0, 0], 120, 31), ([38, 0, 0, 0], 120, 38),
    ([42, 69, 0, 0], 25, 10821), ([69, 42, 0, 0], 25, 17706),
    ([97, 218, 192, 48], 26, 1641726000), ([98, 150, 3, 64], 26, 1653998400),
])
def test_decode_uint(backend, bytes_, v, exp):
    w = calc("decodeUint_test", {"bytes": bytes_, "pos": 0, "v": v}, backend, check_r1cs=True)
    assert w[1] == exp


@pytest.mark.parametrize("pos", [2, 1, 0])
def test_read_type(backend, pos):  # :180-216
    ins = []
    for v in range(256):
        b = [0, 0, 0]
        b[pos] = v
        ins.append({"bytes": b, "pos": pos})
    ws = calc_many("readType_test", ins, backend)
    for v, w in enumerate(ws):
        assert (w[1], w[2], w[3]) == (pos + 1, v >> 5, v)


@pytest.mark.parametrize("circuit,maxlen", [("skipValueScalar_test", 5), ("skipValue5_test", 5)])
def test_skip_value_scalar(backend, circuit, maxlen):  # :219-296
    for n in range(5):
        cb = encodeString("a" * n)
        assert calc(circuit, {"bytes": padArray(cb, maxlen), "pos": 0}, backend)[1] == n + 1
    ws = calc_many(circuit, [{"bytes": padArray(encodeInt(v), maxlen), "pos": 0} for v in range(24)], backend)
    assert [w[1] for w in ws] == [len(encodeInt(v)) for v in range(24)]
    for val in (0xFF, 0xFFFF, 0xFFFFFFFF):
        cb = encodeInt(val)
        assert calc(circuit, {"bytes": padArray(cb, maxlen), "pos": 0}, backend, check_r1cs=True)[1] == len(cb)


@pytest.mark.parametrize("items,circuit,maxlen", [  # :299-364
    ([encodeInt(23)] * 3, "skipValue5_test", 5),
    ([encodeInt(23)] * 4, "skipValue5_test", 5),
    ([encodeInt(0xFF)] * 2, "skipValue5_test", 5),
    ([encodeInt(0xFFFF)], "skipValue5_test", 5),
    ([encodeInt(0xFFFFFFFF)], "skipValue6_test", 6),
    ([encodeString("q"), encodeString("q")], "skipValue5_test", 5),
    ([encodeString("qwe")], "skipValue5_test", 5),
    ([encodeString("q"), encodeInt(0xFF)], "skipValue5_test", 5),
    ([encodeString("q"), encodeInt(23), encodeInt(23)], "skipValue5_test", 5),
])
def test_skip_value_array(backend, items, circuit, maxlen):
    cb = encodeArray(items)
    w = calc(circuit, {"bytes": padArray(cb, maxlen), "pos": 0}, backend, check_r1cs=True)
    assert w[1] == len(cb)


def test_read_string_length(backend):  # :366-380
    for n in range(5):
        w = calc("readStringLength_test", {"bytes": padArray(encodeString("a" * n), 5), "pos": 0}, backend)
        assert (w[1], w[2]) == (n, len(encodeInt(n)))


def test_read_string_length_rejects_non_string(backend):  # hardcore_assert, cbortpl.circom:429
    assert calc("readStringLength_test", {"bytes": padArray(encodeInt(3), 5), "pos": 0}, backend) is None


def test_string_equals(backend):  # :382-404
    s = stringToArray("abcde")
    assert calc("stringEquals_test", {"bytes": padArray(s, 5), "len": 5, "pos": 0}, backend, check_r1cs=True)[1] == 1
    for n in range(6):
        s = stringToArray("b" * n)
        assert calc("stringEquals_test", {"bytes": padArray(s, 5), "len": n, "pos": 0}, backend)[1] == 0


def test_read_map_length(backend):  # :406-429
    e = [(encodeInt(4), encodeInt(5)), (encodeInt(5), encodeInt(4)), (encodeInt(7), encodeInt(3))]
    for n in (1, 2, 3):
        w = calc("readMapLength_test", {"bytes": padArray(encodeMap(e[:n]), 7), "pos": 0}, backend, check_r1cs=True)
        assert w[1] == n
    assert calc("readMapLength_test", {"bytes": padArray(encodeString("ab"), 7), "pos": 0}, backend) is None  # :455


@pytest.mark.parametrize("s", ["", "ab", "abcd"])
def test_copy_string(backend, s):  # :433-476
    w = calc("copyString_test", {"bytes": padArray(encodeString(s), 5), "pos": 0}, backend, check_r1cs=True)
    assert w[1:5] == padArray(stringToArray(s), 4)
    assert (w[5], w[6]) == (len(s) + 1, len(s))
