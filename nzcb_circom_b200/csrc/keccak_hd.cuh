// Streaming Keccak-256 (original 0x01 padding) usable on the device: the Fiat-Shamir transcript of the verifier
// (snarkjs hashToFr over js-sha3 0.8.0 keccak256, un-vendored, /root/reference/yarn.lock:5074).
// keccak.h keeps the host-only one-shot version the prover's host thread uses.
#pragma once
#include <stdint.h>
#include "fp.cuh"

namespace nzcb {

struct KeccakHD {
    uint64_t s[25];
    uint32_t pos;  // bytes absorbed into the current 136-byte block

    NZ_HD void init() {
        for (int i = 0; i < 25; i++) s[i] = 0;
        pos = 0;
    }
    static NZ_HD uint64_t rol(uint64_t x, int n) { return n ? (x << n) | (x >> (64 - n)) : x; }
    NZ_HD void permute() {
        const uint64_t RC[24] = {
            0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL,
            0x000000000000808bULL, 0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL,
            0x000000000000008aULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000aULL,
            0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL, 0x8000000000008003ULL,
            0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800aULL, 0x800000008000000aULL,
            0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
        const int ROT[25] = {0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39, 41, 45, 15, 21, 8, 18, 2, 61, 56, 14};
        for (int r = 0; r < 24; r++) {
            uint64_t C[5], D[5], B[25];
            for (int x = 0; x < 5; x++) C[x] = s[x] ^ s[x + 5] ^ s[x + 10] ^ s[x + 15] ^ s[x + 20];
            for (int x = 0; x < 5; x++) D[x] = C[(x + 4) % 5] ^ rol(C[(x + 1) % 5], 1);
            for (int i = 0; i < 25; i++) s[i] ^= D[i % 5];
            for (int x = 0; x < 5; x++)
                for (int y = 0; y < 5; y++) B[y + 5 * ((2 * x + 3 * y) % 5)] = rol(s[x + 5 * y], ROT[x + 5 * y]);
            for (int x = 0; x < 5; x++)
                for (int y = 0; y < 5; y++) s[x + 5 * y] = B[x + 5 * y] ^ ((~B[(x + 1) % 5 + 5 * y]) & B[(x + 2) % 5 + 5 * y]);
            s[0] ^= RC[r];
        }
    }
    NZ_HD void absorb_byte(uint8_t b) {
        s[pos >> 3] ^= (uint64_t)b << (8 * (pos & 7));
        if (++pos == 136) {
            permute();
            pos = 0;
        }
    }
    NZ_HD void update(const uint8_t* d, uint32_t n) {
        for (uint32_t i = 0; i < n; i++) absorb_byte(d[i]);
    }
    NZ_HD void finish(uint8_t out[32]) {
        s[pos >> 3] ^= (uint64_t)0x01 << (8 * (pos & 7));
        s[16] ^= 0x8000000000000000ULL;  // last byte of the 136-byte rate
        permute();
        for (int i = 0; i < 4; i++)
            for (int k = 0; k < 8; k++) out[8 * i + k] = (uint8_t)(s[i] >> (8 * k));
    }
};

}  // namespace nzcb
