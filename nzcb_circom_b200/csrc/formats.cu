// Host-side format helpers on the drop-in boundary that need no GPU:
//   * circom_runtime's input contract (un-vendored, /root/reference/yarn.lock:2496): signals addressed
//     by the FNV-1a-64 hash of their name, arrays flattened row-major, error codes 1 "Signal not found", 2 "Too many
//     signals set", 3 "Signal already set", 6 "Input signal array access exceeds the size", and the JS wrapper's
//     "Not all inputs have been set" -- what every `cir.calculateWitness(input, true)` of test/nzcp.js:42,
//     test/cbor.js and test/quinSelector.js goes through before a single constraint runs;
//   * `.wtns` v2 writer (snarkjs wtns.calculate / circom_tester's witness array -> file);
//   * `snarkjs zkey export verificationkey` (/root/reference/Makefile:56,61): verification_key.json text from the zkey.
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include <vector>
#include "../../include/nzcb.h"
#include "fp.cuh"

using namespace nzcb;

extern "C" uint64_t nzcb_fnv1a64(const char* name, size_t len) {
    uint64_t h = 0xCBF29CE484222325ull;
    for (size_t i = 0; i < len; i++) {
        h ^= (uint8_t)name[i];
        h *= 0x100000001B3ull;
    }
    return h;
}

// symbol table: "NZSY", u32 version = 1, u32 n_signals, u32 n_inputs_total, then n_signals x {u64 hash, u32 offset, u32 size}
namespace {
struct Sym {
    uint64_t hash;
    uint32_t off, size;
};
int parse_sym(const uint8_t* sym, size_t len, std::vector<Sym>& out, uint32_t& n_in) {
    if (!sym || len < 16 || memcmp(sym, "NZSY", 4) != 0) return NZCB_E_INVALID;
    uint32_t ver, n;
    memcpy(&ver, sym + 4, 4);
    memcpy(&n, sym + 8, 4);
    memcpy(&n_in, sym + 12, 4);
    if (ver != 1 || len != 16 + (size_t)n * 16) return NZCB_E_INVALID;
    out.resize(n);
    for (uint32_t i = 0; i < n; i++) {
        memcpy(&out[i].hash, sym + 16 + (size_t)i * 16, 8);
        memcpy(&out[i].off, sym + 24 + (size_t)i * 16, 4);
        memcpy(&out[i].size, sym + 28 + (size_t)i * 16, 4);
        if ((uint64_t)out[i].off + out[i].size > n_in) return NZCB_E_INVALID;
    }
    return 0;
}
void set_msg(char* err, size_t err_len, const char* fmt, unsigned a = 0, unsigned b = 0) {
    if (err && err_len) snprintf(err, err_len, fmt, a, b);
}
}  // namespace

// The loop of circom_runtime's _doCalculateWitness: for every provided signal k, for every value i:
// setInputSignal(hash, i, value) with the generated runtime's order of checks.  values_le: the provided values of all
// signals back to back, 32 B LE each (any 256-bit value; reduced mod r like Fr.e).  inputs_out: n_inputs x 32 B.
// Returns 0 or -(100 + code) with circom_runtime's message in err; -(100 + 7) = "Not all inputs have been set".
extern "C" int32_t nzcb_inputs_resolve(const uint8_t* sym, size_t sym_len, uint32_t n_signals, const uint64_t* hashes,
                                       const uint32_t* counts, const uint8_t* values_le, uint8_t* inputs_out,
                                       char* err, size_t err_len) {
    std::vector<Sym> tab;
    uint32_t n_in = 0;
    if (parse_sym(sym, sym_len, tab, n_in) != 0 || (n_signals && (!hashes || !counts)) || !inputs_out) {
        set_msg(err, err_len, "bad symbol table or arguments");
        return NZCB_E_INVALID;
    }
    std::vector<uint8_t> is_set(n_in, 0);
    uint32_t remaining = n_in;
    const uint8_t* v = values_le;
    for (uint32_t k = 0; k < n_signals; k++) {
        for (uint32_t i = 0; i < counts[k]; i++, v += 32) {
            if (remaining == 0) {
                set_msg(err, err_len, "Too many signals set");
                return NZCB_E_SIGNAL(2);
            }
            const Sym* s = nullptr;
            for (const Sym& t : tab)
                if (t.hash == hashes[k]) { s = &t; break; }
            if (!s) {
                set_msg(err, err_len, "Signal not found");
                return NZCB_E_SIGNAL(1);
            }
            if (i >= s->size) {
                set_msg(err, err_len, "Input signal array access exceeds the size");
                return NZCB_E_SIGNAL(6);
            }
            if (is_set[s->off + i]) {
                set_msg(err, err_len, "Signal already set");
                return NZCB_E_SIGNAL(3);
            }
            Fr x;
            memcpy(x.v, v, 32);
            x = fr_reduce_256(x);
            memcpy(inputs_out + (size_t)(s->off + i) * 32, x.v, 32);
            is_set[s->off + i] = 1;
            remaining--;
        }
        if (counts[k] == 0) {  // a name with no values still has to exist (getInputSignalSize < 0 -> "Signal not found")
            bool found = false;
            for (const Sym& t : tab) found = found || t.hash == hashes[k];
            if (!found) {
                set_msg(err, err_len, "Signal not found");
                return NZCB_E_SIGNAL(1);
            }
        }
    }
    if (remaining) {
        set_msg(err, err_len, "Not all inputs have been set. Only %u out of %u", n_in - remaining, n_in);
        return NZCB_E_SIGNAL(7);
    }
    return 0;
}

// .wtns v2: magic, version 2, 2 sections; 1 = {n8 = 32, r, nWitness}, 2 = nWitness x 32 B LE canonical
extern "C" int32_t nzcb_wtns_export(const uint8_t* witness_le, uint32_t n_witness, uint8_t* out, size_t* len) {
    if (!len || (n_witness && !witness_le)) return NZCB_E_INVALID;
    const size_t need = 12 + 12 + 40 + 12 + (size_t)n_witness * 32;
    if (!out || *len < need) {
        *len = need;
        return out ? NZCB_E_INVALID : 0;
    }
    uint8_t* p = out;
    auto u32 = [&](uint32_t x) { memcpy(p, &x, 4); p += 4; };
    auto u64 = [&](uint64_t x) { memcpy(p, &x, 8); p += 8; };
    memcpy(p, "wtns", 4); p += 4;
    u32(2); u32(2);
    u32(1); u64(40);
    u32(32);
    for (int i = 0; i < 8; i++) u32(FrParams::mod(i));
    u32(n_witness);
    u32(2); u64((uint64_t)n_witness * 32);
    memcpy(p, witness_le, (size_t)n_witness * 32);
    *len = need;
    return 0;
}

namespace {
std::string limbs_to_dec(const uint32_t* v) {
    uint32_t t[8];
    memcpy(t, v, 32);
    std::string s;
    bool nonzero = true;
    while (nonzero) {
        uint64_t rem = 0;
        nonzero = false;
        for (int k = 7; k >= 0; k--) {
            const uint64_t cur = (rem << 32) | t[k];
            t[k] = (uint32_t)(cur / 1000000000u);
            rem = cur % 1000000000u;
            if (t[k]) nonzero = true;
        }
        char buf[16];
        snprintf(buf, sizeof(buf), nonzero ? "%09u" : "%u", (unsigned)rem);
        s = std::string(buf) + s;
    }
    return s;
}
template <class F>
std::string mont_dec(const uint8_t* lem) {
    F x;
    memcpy(x.v, lem, 32);
    return limbs_to_dec(x.from_mont().v);
}
}  // namespace

// JSON.stringify(vk, null, 1) of snarkjs' zkey export verificationkey: key order protocol, curve, nPublic, power, k1, k2,
// Qm, Ql, Qr, Qo, Qc, S1, S2, S3, X_2, w (snarkjs 0.4.12 src/zkey_export_verificationkey.js, recalled: SURVEY.md A.3)
extern "C" int32_t nzcb_vkey_to_json(const uint8_t* zkey, size_t zlen, char* buf, size_t* len) {
    if (!zkey || !len || zlen < 12 || memcmp(zkey, "zkey", 4) != 0) return NZCB_E_INVALID;
    uint32_t nsec;
    memcpy(&nsec, zkey + 8, 4);
    size_t pos = 12;
    const uint8_t* hdr = nullptr;
    for (uint32_t s = 0; s < nsec && pos + 12 <= zlen; s++) {
        uint32_t id;
        uint64_t size;
        memcpy(&id, zkey + pos, 4);
        memcpy(&size, zkey + pos + 4, 8);
        pos += 12;
        if (size > zlen - pos) return NZCB_E_INVALID;
        if (id == 2) {
            if (size < 156 + 512 + 128) return NZCB_E_INVALID;
            hdr = zkey + pos;
            break;
        }
        pos += size;
    }
    if (!hdr) return NZCB_E_INVALID;
    uint32_t n_public, domain;
    memcpy(&n_public, hdr + 76, 4);
    memcpy(&domain, hdr + 80, 4);
    if (domain == 0 || (domain & (domain - 1))) return NZCB_E_INVALID;
    uint32_t power = 0;
    while ((1u << power) < domain) power++;
    if (power > 28) return NZCB_E_INVALID;
    std::string o = "{\n \"protocol\": \"plonk\",\n \"curve\": \"bn128\",\n";
    o += " \"nPublic\": " + std::to_string(n_public) + ",\n \"power\": " + std::to_string(power) + ",\n";
    o += " \"k1\": \"" + mont_dec<Fr>(hdr + 92) + "\",\n \"k2\": \"" + mont_dec<Fr>(hdr + 124) + "\",\n";
    static const char* names[8] = {"Qm", "Ql", "Qr", "Qo", "Qc", "S1", "S2", "S3"};
    for (int i = 0; i < 8; i++) {
        const uint8_t* g = hdr + 156 + 64 * i;
        bool inf = true;
        for (int k = 0; k < 64; k++) inf = inf && g[k] == 0;
        o += std::string(" \"") + names[i] + "\": [\n  \"" + (inf ? "0" : mont_dec<Fq>(g)) + "\",\n  \"" +
             (inf ? "1" : mont_dec<Fq>(g + 32)) + "\",\n  \"" + (inf ? "0" : "1") + "\"\n ],\n";
    }
    const uint8_t* x2 = hdr + 156 + 512;
    o += " \"X_2\": [\n  [\n   \"" + mont_dec<Fq>(x2) + "\",\n   \"" + mont_dec<Fq>(x2 + 32) + "\"\n  ],\n  [\n   \"" +
         mont_dec<Fq>(x2 + 64) + "\",\n   \"" + mont_dec<Fq>(x2 + 96) + "\"\n  ],\n  [\n   \"1\",\n   \"0\"\n  ]\n ],\n";
    // w = 5^((r-1)/2^28) squared down to the 2^power-th root
    Fr w = Fr::from_u64(5);
    {
        uint32_t e[8], sh[8];
        for (int i = 0; i < 8; i++) e[i] = FrParams::mod(i);
        e[0] -= 1;
        for (int i = 0; i < 8; i++) sh[i] = (e[i] >> 28) | (i + 1 < 8 ? e[i + 1] << 4 : 0);
        w = w.pow_limbs(sh);
    }
    for (uint32_t i = power; i < 28; i++) w = w.sqr();
    o += " \"w\": \"" + limbs_to_dec(w.from_mont().v) + "\"\n}";
    const size_t need = o.size() + 1;
    if (!buf || *len < need) {
        *len = need;
        return buf ? NZCB_E_INVALID : 0;
    }
    memcpy(buf, o.c_str(), need);
    *len = need;
    return 0;
}
