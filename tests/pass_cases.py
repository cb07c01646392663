"""Pass URIs for the ingest tests: the reference's example pass, synthetic passes and malformed variants that hit
every rejection branch of test/helpers/nzcp.js (bad base32, wrong tag, wrong shapes, JS 32-bit length folding)."""
import random

from nzcb_circom_b200 import nzcp_helpers as h

_B32 = "ABCDEFGHIJKLMNOPQRSTUVWXYZ234567"


def b32encode(data: bytes) -> str:
    """unpadded RFC 4648 base32 (what an NZCP QR code carries)"""
    bits = "".join(f"{b:08b}" for b in data)
    bits += "0" * (-len(bits) % 5)
    return "".join(_B32[int(bits[i:i + 5], 2)] for i in range(0, len(bits), 5))


def cose_bytes(protected, payload, sig=bytes(64), head=b"\xd2\x84", unprot=b"\xa0"):
    def bstr(x):
        n = len(x)
        hd = bytes([0x40 + n]) if n <= 23 else bytes([0x58, n]) if n < 256 else bytes([0x59, n >> 8, n & 255])
        return hd + x
    return head + bstr(protected) + unprot + bstr(payload) + bstr(sig)


def synth_uri(seed, live=True):
    p = h.synth_pass(seed, live)
    return p["uri"], p


def cases(n_synth=24, seed=1):
    """[(label, uri)] -- valid and malformed"""
    rng = random.Random(seed)
    out = [("example", h.EXAMPLE_PASS_URI)]
    for i in range(n_synth):
        out.append((f"synth{i}", synth_uri(i, live=(i % 3 != 0))[0]))
    ex = h.EXAMPLE_PASS_URI
    prot, pay = b"\xa2\x04\x45key-1\x01\x26", bytes(rng.randrange(256) for _ in range(200))
    good = cose_bytes(prot, pay)
    u = lambda b: "NZCP:/1/" + b32encode(b)
    out += [
        ("empty", ""), ("short", "NZCP:/"), ("prefix_only", "NZCP:/1/"), ("one_char", "NZCP:/1/A"),
        ("lowercase", ex[:40] + ex[40:].lower()), ("bad_char_1", ex[:100] + "1" + ex[101:]),
        ("bad_char_eq", ex + "="), ("bad_char_last", ex[:-1] + "8"), ("nul_char", ex[:50] + "\x00" + ex[51:]),
        ("other_prefix", "XXXXXXXX" + ex[8:]),  # substring(8): the prefix is not checked
        ("truncated_half", ex[:len(ex) // 2]), ("truncated_1", ex[:-1]), ("truncated_2", ex[:-2]),
        ("truncated_sig", ex[:-100]), ("trailing_junk", ex + "AAAAAAA"),
        ("wrong_tag", u(b"\xd1" + good[1:])), ("no_tag", u(good[1:])),
        ("array3", u(cose_bytes(prot, pay, head=b"\xd2\x83"))), ("array5", u(cose_bytes(prot, pay, head=b"\xd2\x85") + b"\x40")),
        ("array_len_1byte", u(cose_bytes(prot, pay, head=b"\xd2\x98\x04"))),
        ("array_len_2byte", u(cose_bytes(prot, pay, head=b"\xd2\x99\x00\x04"))),
        ("array_len_4byte", u(cose_bytes(prot, pay, head=b"\xd2\x9a\x00\x00\x00\x04"))),
        ("array_len_8byte", u(cose_bytes(prot, pay, head=b"\xd2\x9b\x00\x00\x00\x00\x00\x00\x00\x04"))),
        ("array_len_8byte_fold", u(cose_bytes(prot, pay, head=b"\xd2\x9b\x00\x00\x00\x04\x00\x00\x00\x00"))),  # JS: x << 32 == x << 0
        ("array_len_neg", u(cose_bytes(prot, pay, head=b"\xd2\x9a\x80\x00\x00\x04"))),
        ("array_len_28", u(cose_bytes(prot, pay, head=b"\xd2\x9c"))),
        ("map_not_array", u(cose_bytes(prot, pay, head=b"\xd2\xa4"))),
        ("unprot_empty_array", u(cose_bytes(prot, pay, unprot=b"\x80"))),
        ("unprot_empty_bstr", u(cose_bytes(prot, pay, unprot=b"\x40"))),
        ("unprot_empty_map_long", u(cose_bytes(prot, pay, unprot=b"\xb8\x00"))),
        ("unprot_empty_text", u(cose_bytes(prot, pay, unprot=b"\x60"))),
        ("unprot_int", u(cose_bytes(prot, pay, unprot=b"\x00"))),
        ("unprot_map1", u(cose_bytes(prot, pay, unprot=b"\xa1\x01\x02"))),
        ("unprot_array1", u(cose_bytes(prot, pay, unprot=b"\x81\x01"))),
        ("unprot_tag", u(cose_bytes(prot, pay, unprot=b"\xc0"))),
        ("prot_text", u(b"\xd2\x84\x6a" + prot + b"\xa0\x58\xc8" + pay + b"\x40")),
        ("payload_text", u(b"\xd2\x84\x4a" + prot + b"\xa0\x78\xc8" + pay + b"\x40")),
        ("sig_int", u(b"\xd2\x84\x4a" + prot + b"\xa0\x58\xc8" + pay + b"\x05")),
        ("sig_missing", u(b"\xd2\x84\x4a" + prot + b"\xa0\x58\xc8" + pay)),
        ("sig_overrun", u(b"\xd2\x84\x4a" + prot + b"\xa0\x58\xc8" + pay + b"\x58\x40" + bytes(10))),
        ("sig_len_neg", u(b"\xd2\x84\x4a" + prot + b"\xa0\x58\xc8" + pay + b"\x5a\xff\xff\xff\xff")),
        ("empty_fields", u(b"\xd2\x84\x40\xa0\x40\x40")),
        ("payload_300", u(cose_bytes(prot, bytes(rng.randrange(256) for _ in range(300))))),
        ("payload_800", u(cose_bytes(prot, bytes(rng.randrange(256) for _ in range(800))))),
        ("prot_30", u(cose_bytes(bytes(range(30)), pay))),
        ("payload_len_2byte_small", u(b"\xd2\x84\x4a" + prot + b"\xa0\x59\x00\xc8" + pay + b"\x40")),
        # 5n % 8 != 0: the parser runs into the zero tail byte of the ceil-sized Uint8Array (a uint 0, not a bstr)
        ("tail_byte_read", u(b"\xd2\x84\x4a" + prot + b"\xa0\x58\xc8" + pay)[:-1] + "A"),
        ("too_long", "NZCP:/1/" + "A" * 5000),
    ]
    return out
