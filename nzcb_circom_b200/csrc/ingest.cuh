// Pass ingest arithmetic shared by the kernel (ingest.cu) and the host-side check (tests/hostcheck): base32 symbols,
// the COSE_Sign1 field walk with JavaScript's number semantics, ToBeSigned bytes and the marshalled input values.
// Restates /root/reference/test/helpers/nzcp.js:9-24,26-56,58-105,123-137,141-172,180-206 and
// /root/reference/test/helpers/utils.js:2,49,71,87 (see ingest.cu).
#pragma once
#include <stdint.h>
#include "fp.cuh"

namespace nzcb {
constexpr uint32_t ING_MAX_CHARS = 4096;                  // longest pass URI accepted (a QR code holds < 3 k)
constexpr uint32_t ING_MAX_RAW = ING_MAX_CHARS * 5 / 8 + 1;
constexpr uint32_t ING_MAX_TBS = 1024;                    // largest maxLen (the circuits use 314 and 351)
constexpr uint32_t ING_REJECT_LEN = 0xFFFF;               // toBeSignedLen of an undecodable pass: no circuit accepts it

NZ_HD int b32_val(uint8_t ch) {
    if (ch >= 'A' && ch <= 'Z') return ch - 'A';
    if (ch >= '2' && ch <= '7') return ch - '2' + 26;
    return -1;
}

// byte j of base32ToBytes(sym[0..n)) (nzcp.js:9-24): stream bits [8j, 8j+8) live in symbols 8j/5 .. (8j+7)/5; the
// Uint8Array has ceil(5n/8) entries, so a last partial byte exists and stays zero
NZ_HD uint8_t b32_out_byte(const uint8_t* sym, uint32_t n, uint32_t j) {
    if (j >= n * 5 / 8) return 0;
    const uint32_t s0 = 8 * j / 5;
    uint32_t acc = 0;
    for (uint32_t k = 0; k < 3; k++) acc = (acc << 5) | (s0 + k < n ? (uint32_t)b32_val(sym[s0 + k]) : 0u);
    return (uint8_t)(acc >> (7 - (8 * j - 5 * s0)));
}

// Stream of nzcp.js:26-56
struct Rd {
    const uint8_t* d;
    uint32_t ptr, len;
    bool bad;
    NZ_HD uint32_t getc() {
        if (ptr >= len) {
            bad = true;
            return 0;
        }
        return d[ptr++];
    }
};

// decodeUint of nzcp.js:60-86 with JavaScript's 32-bit shift semantics; a negative result is returned as such
NZ_HD int32_t cbor_uint(Rd& s, uint32_t v) {
    uint32_t x = v & 31;
    if (x <= 23) return (int32_t)x;
    if (x == 24) return (int32_t)s.getc();
    if (x == 25) {
        x = s.getc() << 8;
        return (int32_t)(x | s.getc());
    }
    if (x == 26 || x == 27) {
        uint32_t r = 0;
        const int nb = x == 26 ? 4 : 8;
        for (int i = 0; i < nb; i++) r |= s.getc() << ((8 * (nb - 1 - i)) & 31);
        return (int32_t)r;
    }
    s.bad = true;
    return 0;
}

// a CBOR byte string at the cursor: its offset and length (chop of nzcp.js:44-55)
NZ_HD bool cbor_bstr(Rd& s, uint32_t* off, uint32_t* len) {
    const uint32_t v = s.getc();
    if (s.bad || (v >> 5) != 2) return false;
    const int32_t n = cbor_uint(s, v);
    if (s.bad || n < 0 || (uint64_t)s.ptr + (uint32_t)n > s.len) return false;
    *off = s.ptr;
    *len = (uint32_t)n;
    s.ptr += (uint32_t)n;
    return true;
}

struct CoseFields {
    uint32_t ok, prot_off, prot_len, pay_off, pay_len, tbs_len, hdr_prot, hdr_pay;
};

NZ_HD uint32_t enc_hdr_len(uint32_t n) { return n <= 23 ? 1 : n < 256 ? 2 : 3; }
// byte i of encodeBytes(data) (nzcp.js:123-137)
NZ_HD uint8_t enc_bytes_at(const uint8_t* data, uint32_t n, uint32_t hdr, uint32_t i) {
    if (i >= hdr) return data[i - hdr];
    if (hdr == 1) return (uint8_t)(0x40 + n);
    if (hdr == 2) return i == 0 ? 0x58 : (uint8_t)n;
    return i == 0 ? 0x59 : i == 1 ? (uint8_t)(n >> 8) : (uint8_t)n;
}

// decodeCOSE (nzcp.js:152-172): tag 0xd2, an array of 4 = [bstr, empty object, bstr, bstr].  decodeCBORStream
// decodes every item before decodeCOSE inspects them, but any item other than these four shapes is rejected either
// by the decoder or by the type check, so the flat walk accepts exactly the same byte strings.
NZ_HD CoseFields parse_cose(const uint8_t* raw, uint32_t n_raw) {
    CoseFields g = {};
    Rd s{raw, 0, n_raw, false};
    bool ok = s.getc() == 0xd2 && !s.bad;
    uint32_t v = 0;
    if (ok) {
        v = s.getc();
        ok = !s.bad && (v >> 5) == 4;
    }
    if (ok) ok = cbor_uint(s, v) == 4 && !s.bad;
    if (ok) ok = cbor_bstr(s, &g.prot_off, &g.prot_len);
    if (ok) {
        // data[1]: typeof 'object' with no own keys -- an empty map, array or byte string
        v = s.getc();
        const uint32_t t = v >> 5;
        ok = !s.bad && (t == 5 || t == 4 || t == 2);
        if (ok) ok = cbor_uint(s, v) == 0 && !s.bad;
    }
    if (ok) ok = cbor_bstr(s, &g.pay_off, &g.pay_len);
    uint32_t so = 0, sl = 0;
    if (ok) ok = cbor_bstr(s, &so, &sl);
    if (ok) {
        g.hdr_prot = enc_hdr_len(g.prot_len);
        g.hdr_pay = enc_hdr_len(g.pay_len);
        g.tbs_len = 12 + g.hdr_prot + g.prot_len + 1 + g.hdr_pay + g.pay_len;
        g.ok = 1;
    }
    return g;
}

// byte t of fitBytes(encodeToBeSigned(bodyProtected, payload), maxLen) (nzcp.js:180-206, utils.js:49)
NZ_HD uint8_t tbs_byte_at(const CoseFields& g, const uint8_t* raw, uint32_t t) {
    if (!g.ok || t >= g.tbs_len) return 0;
    const uint32_t a = 12, c = a + g.hdr_prot + g.prot_len;
    if (t < a) {  // array(4), text(10) "Signature1"
        const uint8_t head[12] = {0x84, 0x6a, 'S', 'i', 'g', 'n', 'a', 't', 'u', 'r', 'e', '1'};
        return head[t];
    }
    if (t < c) return enc_bytes_at(raw + g.prot_off, g.prot_len, g.hdr_prot, t - a);
    if (t == c) return 0x40;  // external_aad: empty
    return enc_bytes_at(raw + g.pay_off, g.pay_len, g.hdr_pay, t - c - 1);
}

// main input i of NZCPPubIdentity (nzcptpl.circom:486-488) as test/nzcp.js:36-41 builds it: toBeSigned bits MSB
// first per byte, the TRUE toBeSignedLen, bufferToBitArray(evmRearrangeBytes(data)) -- position (19-k)*8 + p is
// bit p of data byte k
NZ_HD uint32_t ingest_input_value(const CoseFields& g, const uint8_t* tbs_fitted, const uint8_t* dat20,
                                  uint32_t max_len, uint32_t i) {
    if (i < 8 * max_len) return (tbs_fitted[i >> 3] >> (7 - (i & 7))) & 1;
    if (i == 8 * max_len) return g.ok ? g.tbs_len : ING_REJECT_LEN;
    if (!g.ok) return 0;
    const uint32_t q = i - 8 * max_len - 1;
    return (dat20[19 - (q >> 3)] >> (q & 7)) & 1;
}
}  // namespace nzcb
