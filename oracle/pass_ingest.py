"""TEST INFRASTRUCTURE -- CPU oracle of the pass ingest path (QR string -> circuit inputs).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this; the product path is
csrc/ingest.cu (nzcb_pass_ingest_batch).

Restates, statement by statement and with JavaScript's number semantics (32-bit shifts), what the reference's
tests do to a pass before `calculateWitness`:

  /root/reference/test/helpers/nzcp.js:9-24    base32ToBytes  (output array of ceil(5n/8) bytes: a zero tail byte)
  /root/reference/test/helpers/nzcp.js:26-56   Stream.getc / chop
  /root/reference/test/helpers/nzcp.js:58-105  decodeCBORStream (decodeUint: `<<` is a 32-bit shift in JS, so the
                                               64-bit form folds onto 32 bits and a 32-bit length can go negative)
  /root/reference/test/helpers/nzcp.js:141-172 decodeBytes (substring(8), prefix unchecked), decodeCOSE
  /root/reference/test/helpers/nzcp.js:123-137 encodeBytes, :180-206 encodeToBeSigned
  /root/reference/test/helpers/utils.js:2,49,71,87   bufferToBitArray, fitBytes, evmRearrangeBits/Bytes
  /root/reference/test/nzcp.js:36-41           the input object {toBeSigned, toBeSignedLen, data}

Pinned by the reference's own vectors: EXAMPLE_PASS_URI (test/nzcp.js:71) -> ToBeSigned whose SHA-256 is
test/utils.js:17 `example2` (tests/test_pass_ingest.py).
"""

_B32 = "ABCDEFGHIJKLMNOPQRSTUVWXYZ234567"


class InvalidData(Exception):
    pass


def _i32(x):
    x &= 0xFFFFFFFF
    return x - (1 << 32) if x & 0x80000000 else x


def base32ToBytes(s):
    n = len(s)
    out = bytearray((n * 5 + 7) // 8)
    buff = bits = outp = 0
    for ch in s:
        val = _B32.find(ch) if len(ch) == 1 else -1
        if val < 0:
            raise InvalidData("invalid data")
        buff = _i32((buff << 5) | val)
        bits += 5
        if bits >= 8:
            bits -= 8
            out[outp] = (buff >> bits) & 0xFF
            outp += 1
    return bytes(out)


class Stream:
    def __init__(self, data):
        self.data, self.ptr, self.len = data, 0, len(data)

    def getc(self):
        if self.ptr >= self.len:
            raise InvalidData("invalid data")
        self.ptr += 1
        return self.data[self.ptr - 1]

    def chop(self, n):
        if n < 0:
            raise InvalidData("invalid length")
        if self.ptr + n > self.len:
            raise InvalidData("invalid data")
        self.ptr += n
        return self.data[self.ptr - n:self.ptr]


def decodeUint(stream, v):
    x = v & 31
    if x <= 23:
        return x
    if x == 24:
        return stream.getc()
    if x == 25:
        x = stream.getc() << 8
        return x | stream.getc()
    if x == 26:
        x = _i32(stream.getc() << 24)
        for sh in (16, 8, 0):
            x = _i32(x | (stream.getc() << sh))
        return x
    if x == 27:
        x = 0
        for sh in (56, 48, 40, 32, 24, 16, 8, 0):
            x = _i32(x | _i32(stream.getc() << (sh & 31)))
        return x
    raise InvalidData("invalid data")


class _Bytes(bytes):
    """marks a CBOR byte string (JS: Uint8Array) apart from a text string"""


def decodeCBORStream(stream, depth=0):
    if depth > 900:  # JS: "Maximum call stack size exceeded" -- a throw either way
        raise InvalidData("invalid data")
    v = stream.getc()
    t = v >> 5
    if t == 0:
        return decodeUint(stream, v)
    if t == 1:
        return ~decodeUint(stream, v)
    if t == 2:
        return _Bytes(stream.chop(decodeUint(stream, v)))
    if t == 3:
        return bytes(stream.chop(decodeUint(stream, v))).decode("utf-8", "replace")
    if t in (4, 5):
        n = decodeUint(stream, v)
        # new Array(n): RangeError when negative; every element needs at least one byte of the stream
        if n < 0 or n * (t - 3) > stream.len - stream.ptr:
            raise InvalidData("invalid data")
        if t == 4:
            return [decodeCBORStream(stream, depth + 1) for _ in range(n)]
        return [(decodeCBORStream(stream, depth + 1), decodeCBORStream(stream, depth + 1)) for _ in range(n)]
    raise InvalidData("This QR code is invalid.")


def _is_empty_object(x):
    """typeof x === 'object' && Object.keys(x).length === 0: {} (map as list of pairs here), [] or an empty
    Uint8Array"""
    if isinstance(x, _Bytes):
        return len(x) == 0
    if isinstance(x, list):
        return len(x) == 0
    return False


def decodeCOSE(data):
    st = Stream(data)
    if st.getc() != 0xD2:
        raise InvalidData("invalid data")
    first = st.data[st.ptr] if st.ptr < st.len else None
    d = decodeCBORStream(st)
    is_array = first is not None and (first >> 5) == 4
    if not (is_array and isinstance(d, list) and len(d) == 4 and isinstance(d[0], _Bytes) and _is_empty_object(d[1])
            and isinstance(d[2], _Bytes) and isinstance(d[3], _Bytes)):
        raise InvalidData("invalid data")
    return {"bodyProtected": bytes(d[0]), "payload": bytes(d[2]), "signature": bytes(d[3])}


def encodeBytes(data):
    x = len(data)
    if x <= 23:
        return bytes([0x40 + x]) + bytes(data)
    if x < 256:
        return bytes([0x58, x]) + bytes(data)
    if x < 65536:
        return bytes([0x59, x >> 8, x & 0xFF]) + bytes(data)
    raise InvalidData("Too big data")


def encodeToBeSigned(bodyProtected, payload):
    return b"\x84\x6aSignature1" + encodeBytes(bodyProtected) + encodeBytes(b"") + encodeBytes(payload)


def to_be_signed(passURI):
    """getCOSE + encodeToBeSigned of one pass URI (a str); raises InvalidData"""
    cose = decodeCOSE(base32ToBytes(passURI[8:]))
    return encodeToBeSigned(cose["bodyProtected"], cose["payload"])


def circuit_inputs(toBeSigned, maxLen, data20):
    """flattened main inputs in declaration order (nzcptpl.circom:486-488): toBeSigned[8*maxLen] bits MSB first
    per byte of fitBytes(toBeSigned, maxLen), toBeSignedLen (the TRUE length), data[160] =
    bufferToBitArray(evmRearrangeBytes(data20))"""
    fitted = bytearray(maxLen)
    for i, b in enumerate(toBeSigned):
        if i < maxLen:  # a Uint8Array ignores out-of-range writes
            fitted[i] = b
    bits = [(fitted[i >> 3] >> (7 - (i & 7))) & 1 for i in range(8 * maxLen)]
    n = len(data20)
    dbits = [0] * (8 * n)
    for k in range(n):
        for i in range(8):  # evmRearrangeBits: res[(n-1-k)*8 + 7-i] = MSB-first bit i of byte k
            dbits[(n - 1 - k) * 8 + (7 - i)] = (data20[k] >> (7 - i)) & 1
    return bits + [len(toBeSigned)] + dbits


REJECT_LEN = 0xFFFF  # what the device puts into toBeSignedLen of a pass it could not decode


def ingest(passURI, maxLen, data20):
    """-> (status, fitted toBeSigned bytes[maxLen], true length, inputs list) as nzcb_pass_ingest_batch returns them"""
    try:
        tbs = to_be_signed(passURI)
    except InvalidData:
        return -1, bytes(maxLen), 0, [0] * (8 * maxLen) + [REJECT_LEN] + [0] * 160
    fitted = (tbs + bytes(maxLen))[:maxLen]
    return 0, fitted, len(tbs), circuit_inputs(tbs, maxLen, data20)
