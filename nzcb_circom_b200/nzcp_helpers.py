"""Host-side input marshalling: Python mirrors of the reference's JS test helpers
(same names, same argument meaning), plus the synthetic pass generator.

  test/helpers/utils.js   bufferToBitArray :2, bitArrayToBuffer :12, chunkToBits :40,
                          fitBytes :49, bitArrayToNum :57, evmRearrangeBits :71,
                          evmBytesToNum :83, evmRearrangeBytes :87
  test/helpers/nzcp.js    base32ToBytes :9-24, decodeCBORStream :58-105, decodeCOSE :152-174,
                          getCOSE :176, encodeToBeSigned :180-206
  test/helpers/cbor.js    encodeUint :10, encodeInt :28, encodeString :33, encodeArray :42,
                          encodeMap :47, padArray :52
"""
import random

# ---- test/helpers/utils.js -------------------------------------------------


def bufferToBitArray(b):
    return [(x >> (7 - j)) & 1 for x in b for j in range(8)]


def bitArrayToBuffer(a):
    out = bytearray((len(a) - 1) // 8 + 1 if a else 0)
    for i, bit in enumerate(a):
        out[i // 8] |= int(bit) << (7 - (i % 8))
    return bytes(out)


def bufferToBytes(b):
    return list(b)


def chunkToBits(chunk, chunkSize):
    return [(int(chunk) >> j) & 1 for j in range(chunkSize)]


def fitBytes(data, maxLen):
    out = bytearray(maxLen)
    out[:len(data)] = data
    return bytes(out)


def bitArrayToNum(a):
    num = 0
    for i, bit in enumerate(a):
        num |= int(bit) << i
    return num


def evmRearrangeBits(bitArray):
    n = len(bitArray) // 8
    res = [0] * len(bitArray)
    for k in range(n):
        b = n - 1 - k
        for i in range(8):
            res[b * 8 + (7 - i)] = bitArray[k * 8 + i]
    return res


def evmBytesToNum(b):
    return bitArrayToNum(bufferToBitArray(bitArrayToBuffer(evmRearrangeBits(bufferToBitArray(b)))))


def evmRearrangeBytes(b):
    return bitArrayToBuffer(evmRearrangeBits(bufferToBitArray(b)))


# ---- test/helpers/nzcp.js --------------------------------------------------
_B32 = "ABCDEFGHIJKLMNOPQRSTUVWXYZ234567"


def base32ToBytes(s):
    out = bytearray()
    buff = bits = 0
    for ch in s:
        val = _B32.find(ch)
        if val < 0:
            raise ValueError("invalid data")
        buff = (buff << 5) | val
        bits += 5
        if bits >= 8:
            bits -= 8
            out.append((buff >> bits) & 0xFF)
    return bytes(out)


class _Stream:
    def __init__(self, data):
        self.data, self.ptr = data, 0

    def getc(self):
        if self.ptr >= len(self.data):
            raise ValueError("invalid data")
        self.ptr += 1
        return self.data[self.ptr - 1]

    def chop(self, n):
        if n < 0 or self.ptr + n > len(self.data):
            raise ValueError("invalid data")
        self.ptr += n
        return self.data[self.ptr - n:self.ptr]


def decodeCBORStream(stream):
    def uint(v):
        x = v & 31
        if x <= 23:
            return x
        if x in (24, 25, 26, 27):
            n = 1 << (x - 24)
            return int.from_bytes(bytes(stream.getc() for _ in range(n)), "big")
        raise ValueError("invalid data")

    def dec():
        v = stream.getc()
        t = v >> 5
        if t == 0:
            return uint(v)
        if t == 1:
            return ~uint(v)
        if t == 2:
            return bytes(stream.chop(uint(v)))
        if t == 3:
            return bytes(stream.chop(uint(v))).decode("utf-8")
        if t == 4:
            return [dec() for _ in range(uint(v))]
        if t == 5:
            return {dec(): dec() for _ in range(uint(v))}
        raise ValueError("This QR code is invalid.")

    return dec()


def decodeBytes(passURI):
    return base32ToBytes(passURI[8:])


def decodeCOSE(data):
    st = _Stream(data)
    if st.getc() != 0xD2:
        raise ValueError("invalid data")
    d = decodeCBORStream(st)
    if not (isinstance(d, list) and len(d) == 4 and isinstance(d[0], bytes) and d[1] == {} and isinstance(d[2], bytes)
            and isinstance(d[3], bytes)):
        raise ValueError("invalid data")
    return {"bodyProtected": d[0], "payload": d[2], "signature": d[3]}


def getCOSE(passURI):
    return decodeCOSE(decodeBytes(passURI))


def _encodeBytes(data):
    x = len(data)
    if x <= 23:
        return bytes([0x40 + x]) + bytes(data)
    if x < 256:
        return bytes([0x58, x]) + bytes(data)
    if x < 65536:
        return bytes([0x59, x >> 8, x & 0xFF]) + bytes(data)
    raise ValueError("Too big data")


def encodeToBeSigned(bodyProtected, payload):
    """COSE Sig_structure: [ "Signature1", body_protected, external_aad = h'', payload ]"""
    return b"\x84\x6aSignature1" + _encodeBytes(bodyProtected) + _encodeBytes(b"") + _encodeBytes(payload)


# ---- test/helpers/cbor.js --------------------------------------------------
def encodeUint(val):
    if val <= 23:
        return [val]
    if val <= 0xFF:
        return [24, val]
    if val <= 0xFFFF:
        return [25, val >> 8, val & 0xFF]
    if val <= 0xFFFFFFFF:
        return [26, val >> 24, (val >> 16) & 0xFF, (val >> 8) & 0xFF, val & 0xFF]
    raise ValueError("Value too large")


def encodeInt(val):
    x, *rest = encodeUint(val)
    return [(0 << 5) | x, *rest]


def stringToArray(s):
    return [ord(ch) for ch in s]


def encodeString(s):
    x, *rest = encodeUint(len(s))
    return [(3 << 5) | x, *rest, *stringToArray(s)]


def encodeArray(arr):
    x, *rest = encodeUint(len(arr))
    return [(4 << 5) | x, *rest, *[b for item in arr for b in item]]


def encodeMap(entries):
    """entries: list of (encoded key bytes, encoded value bytes).  The JS builds an object keyed by
    the stringified key bytes and re-parses them with parseInt -- for the single-byte int keys the
    tests use that is the key byte itself."""
    x, *rest = encodeUint(len(entries))
    out = [(5 << 5) | x, *rest]
    for k, v in entries:
        out += list(k) + list(v)
    return out


def padArray(arr, n):
    return list(arr) + [0] * max(n - len(arr), 0)


# ---- the reference's embedded example pass (test/nzcp.js:71) ---------------
EXAMPLE_PASS_URI = ("NZCP:/1/2KCEVIQEIVVWK6JNGEASNICZAEP2KALYDZSGSZB2O5SWEOTOPJRXALTDN53GSZBRHEXGQZLBNR2GQLTOPICRUYMBTIFAIGTUKBAAUYTW"
                    "MOSGQQDDN5XHIZLYOSBHQJTIOR2HA4Z2F4XXO53XFZ3TGLTPOJTS6MRQGE4C6Y3SMVSGK3TUNFQWY4ZPOYYXQKTIOR2HA4Z2F4XW46TDOAXGG33W"
                    "NFSDCOJONBSWC3DUNAXG46RPMNXW45DFPB2HGL3WGFTXMZLSONUW63TFGEXDALRQMR2HS4DFQJ2FMZLSNFTGSYLCNRSUG4TFMRSW45DJMFWG6UDV"
                    "MJWGSY2DN53GSZCQMFZXG4LDOJSWIZLOORUWC3CTOVRGUZLDOSRWSZ3JOZSW4TTBNVSWISTBMNVWUZTBNVUWY6KOMFWWKZ2TOBQXE4TPO5RWI33C"
                    "NIYTSNRQFUYDILJRGYDVAYFE6VGU4MCDGK7DHLLYWHVPUS2YIDJOA6Y524TD3AZRM263WTY2BE4DPKIF27WKF3UDNNVSVWRDYIYVJ65IRJJJ6Z25"
                    "M2DO4YZLBHWFQGVQR5ZLIWEQJOZTS3IQ7JTNCFDX")
EXAMPLE_TOBESIGNED_MAX = 314
LIVE_TOBESIGNED_MAX = 351


def nzcp_input(toBeSigned: bytes, maxLen: int, origData: bytes = bytes(range(1, 21))):
    """the input object of testNZCPPubIdentity (test/nzcp.js:36-41)"""
    return {"toBeSigned": bufferToBitArray(fitBytes(toBeSigned, maxLen)), "toBeSignedLen": len(toBeSigned),
            "data": bufferToBitArray(evmRearrangeBytes(origData))}


def nzcp_decode_outputs(out):
    """out[0..2] -> (nullifierHashPart 32 B, toBeSignedHash 32 B, exp, data 20 B)  (test/nzcp.js:44-68)"""
    o = [bitArrayToBuffer(evmRearrangeBits(chunkToBits(x, 248))) for x in out]
    nullifier_hash_part = o[0] + o[1][:1]
    tbs_hash = o[1][1:] + o[2][:2]
    exp = evmBytesToNum(o[2][2:6])
    return nullifier_hash_part, tbs_hash, exp, o[2][6:26]


def bytesToBase32(data):
    """unpadded RFC 4648 base32, the inverse of base32ToBytes (what an NZCP QR code carries)"""
    bits = "".join(f"{x:08b}" for x in data)
    bits += "0" * (-len(bits) % 5)
    return "".join(_B32[int(bits[i:i + 5], 2)] for i in range(0, len(bits), 5))


def passURI(bodyProtected, payload, signature):
    """COSE_Sign1 (tag 18, [protected, {}, payload, signature]) as an "NZCP:/1/" URI: what getCOSE parses"""
    return "NZCP:/1/" + bytesToBase32(b"\xd2\x84" + _encodeBytes(bodyProtected) + b"\xa0" + _encodeBytes(payload)
                                      + _encodeBytes(signature))


# ---- synthetic passes (no network, no real passes) -----------
# bytes 76..246 of the example ToBeSigned: the fixed 171-byte "vc" map prefix up to and including the
# credentialSubject map header (a3); CREDENTIAL_SUBJECT_VC_OFFSET = 171 (nzcptpl.circom:461)
_VC_PREFIX = bytes.fromhex(
    "a46840636f6e7465787482782668747470733a2f2f7777772e77332e6f72672f323031382f63726564656e7469616c732f7631782a"
    "68747470733a2f2f6e7a63702e636f76696431392e6865616c74682e6e7a2f636f6e74657874732f76316776657273696f6e65312e"
    "302e306474797065827456657269666961626c6543726564656e7469616c6f5075626c6963436f766964506173737163726564656e"
    "7469616c5375626a656374a3")
assert len(_VC_PREFIX) == 171
_LETTERS = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz"


def synth_pass(seed, live=True):
    """A random pass that satisfies the circuit's hard-coded offsets (claims map at 30 live / 27
    example, vc prefix of 171 bytes, credential subject of three text entries).  Returns a dict
    with toBeSigned bytes, the expected nullifier / exp, and the 20-byte pass-through data."""
    rng = random.Random(0x6E7A6362 + seed)
    max_len = LIVE_TOBESIGNED_MAX if live else EXAMPLE_TOBESIGNED_MAX
    while True:
        lg = rng.randint(1, 21)
        lf = rng.randint(1, 21)
        given = "".join(rng.choice(_LETTERS) for _ in range(lg))
        family = "".join(rng.choice(_LETTERS) for _ in range(lf))
        dob = f"{rng.randint(1900, 2021):04d}-{rng.randint(1, 12):02d}-{rng.randint(1, 28):02d}"
        nbf = rng.randrange(1 << 30, 1 << 32)
        exp = rng.randrange(1 << 30, 1 << 32)
        if live:
            protected = b"\xa2\x04\x48" + bytes(rng.randrange(256) for _ in range(8)) + b"\x01\x26"
            iss = "did:web:nzcp.identity.health.nz"
        else:
            protected = b"\xa2\x04\x45key-1\x01\x26"
            iss = "did:web:nzcp.covid19.health.nz"
        entries = [("givenName", given), ("familyName", family), ("dob", dob)]
        rng.shuffle(entries)
        cs = b"".join(bytes(encodeString(k)) + bytes(encodeString(v)) for k, v in entries)
        payload = (b"\xa5" + b"\x01" + bytes(encodeString(iss)) + b"\x05" + bytes(encodeInt(nbf)) + b"\x04"
                   + bytes(encodeInt(exp)) + bytes(encodeString("vc")) + _VC_PREFIX + cs + b"\x07\x50"
                   + bytes(rng.randrange(256) for _ in range(16)))
        tbs = encodeToBeSigned(protected, payload)
        if len(tbs) > max_len:
            continue
        # over-read hazard: every CopyString / StringEquals read must stay inside the buffer
        cs_start = tbs.index(_VC_PREFIX) + 171
        pos = cs_start
        ok = True
        for k, v in entries:
            key_first = pos + 1
            val_first = pos + 1 + len(k) + 1
            if key_first + 9 > max_len - 1 or val_first + 20 > max_len - 1:
                ok = False
            pos = val_first + len(v)
        if not ok:
            continue
        data = bytes(rng.randrange(256) for _ in range(20))
        sig = bytes(rng.randrange(256) for _ in range(64))  # never checked by the circuit (the contract does that)
        return {"toBeSigned": tbs, "nullifier": f"{given},{family},{dob}", "exp": exp, "nbf": nbf, "data": data,
                "maxLen": max_len, "uri": passURI(protected, payload, sig)}
