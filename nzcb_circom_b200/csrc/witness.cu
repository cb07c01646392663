// Batched witness calculation -- replaces the circom-generated WASM +
// circom_runtime 0.1.17 WitnessCalculator (un-vendored,
// /root/reference/yarn.lock:2496; call sites /root/reference/test/nzcp.js:42 and
// the 72 other calculateWitness calls of test/cbor.js, test/quinSelector.js,
// test/nzcp.js).  The circuit arrives as a witness program (.wprog, emitted by
// nzcb_circom_b200/circom/builder.py): one instruction per wire, sorted by
// dependency level.
//
// Execution model: one CTA (256 threads) per pass; a level's instructions are spread over the
// CTA -- long linear combinations one per warp, the rest one per thread from records prefetched
// into shared memory -- and levels are separated by __syncthreads.  Wires live in HBM in
// canonical (non-Montgomery) little-endian form, pass-major, so a pass's first
// nWitness wires ARE the payload of its .wtns file and feed the prover without a
// copy.  Constants are kept in both forms, so const * wire is at most a single multiply
// (none for bit wires); wire * wire takes two unless both fit 32 bits.  IsZero's inverse hint
// hits a 2 x 1024-entry table for the |x| <= 1024 operands the CBOR selectors produce
//  and falls back to a Fermat inverse otherwise.
#include <stdlib.h>
#include "common.cuh"
#include <algorithm>
#include <cuda_pipeline.h>

using namespace nzcb;

enum { OP_LIN = 1, OP_MUL = 2, OP_BITS = 3, OP_INV = 4, OP_ASSERT = 5, OP_BITSLC = 6, OP_SHAROUND = 7, OP_SHASCHED = 8, OP_QUINSEL = 9, OP_SHABLOCK = 10, OP_SHAROUNDS = 11 };
constexpr uint32_t NZ_INV_TAB = 1024;
constexpr uint32_t NZ_LONG_LC = 24;  // an LC (or bit decomposition) longer than this is evaluated by a whole warp

struct nzcb_circuit {
    nzcb_ctx* ctx = nullptr;
    uint32_t n_total = 0, n_witness = 0, n_out = 0, n_in = 0, n_consts = 0, n_instr = 0, n_levels = 0, n_code = 0;
    Fr* d_consts = nullptr;      // Montgomery
    Fr* d_consts_can = nullptr;  // canonical
    uint32_t* d_ioff = nullptr;  // instruction offsets; inside a level the long instructions come first
    uint32_t* d_lstart = nullptr;
    uint32_t* d_nlong = nullptr; // per level: how many leading instructions run one-per-warp
    uint4* d_rec = nullptr;      // 128-byte records: a short instruction's whole encoding in one line (prefetched to smem)
    uint32_t* d_code = nullptr;
    Fr* d_invtab = nullptr;      // canonical 1/k, k = 0..NZ_INV_TAB (entry 0 unused)
};

namespace {

struct ProgView {
    const Fr* consts;      // Montgomery
    const Fr* consts_can;  // canonical
    const uint32_t *ioff, *lstart, *nlong, *code;
    const uint4* rec;      // 8 x uint4 (128 B) per instruction, level order: the first 32 code words of each instruction
    const Fr* invtab;
    uint32_t n_total, n_out, n_in, n_levels;
};

__device__ __forceinline__ bool fits_u32(const Fr& v) {
    uint32_t hi = 0;
#pragma unroll
    for (int i = 1; i < 8; i++) hi |= v.v[i];
    return hi == 0;
}

// the multiply and the Fermat inverse of the rare paths (general coefficients, wide operands, operands outside
// the inverse table) stay out of line: the interpreter is latency bound and its code should stay small
__device__ __noinline__ Fr wit_mul(const Fr& a, const Fr& b) { return a * b; }
__device__ __noinline__ Fr wit_inv(const Fr& v) { return v.to_mont().inv().from_mont(); }

// acc += coef_c * v  for one LC term (canonical values; coefficient index c: 0 = +1, 1 = -1)
__device__ __forceinline__ void lc_term(const ProgView& pv, Fr& acc, uint32_t c, const Fr& v) {
    if (c == 0) acc = acc + v;
    else if (c == 1) acc = acc - v;
    else if (v.is_zero()) return;
    else if (fits_u32(v) && v.v[0] == 1) acc = acc + pv.consts_can[c];  // bit wires: no multiply
    else acc = acc + wit_mul(pv.consts[c], v);  // Montgomery const x canonical wire = canonical
}

// CTA sizes: 384 threads (one CTA fills an SM's registers) gives the shortest single-pass latency; batches larger
// than the SM count run 128-thread CTAs, three per SM, so that one pass's barrier waits overlap another's work
constexpr uint32_t WIT_THREADS = 384;
constexpr uint32_t WIT_THREADS_BATCH = 128;
constexpr uint32_t REC_WORDS = 32;
constexpr uint32_t REC_LONG = 0x100u;  // flag on word 0: the encoding does not fit a record, word 1 = its code offset

// where an instruction's words come from: the code stream in global memory, or this thread's record in shared
// memory (word-major so that the threads of a warp read consecutive 16-byte groups)
struct GlobalCode {
    const uint32_t* code;
    __device__ __forceinline__ uint32_t operator()(uint32_t p) const { return code[p]; }
};
template <uint32_t T>
struct SmemCode {
    const uint32_t* base;  // &slot[0][thread] viewed as words
    __device__ __forceinline__ uint32_t operator()(uint32_t k) const { return base[(k >> 2) * (T * 4) + (k & 3)]; }
};

// canonical value of  k + sum coef_i * w_i ;  advances p past the encoded LC   (one thread)
template <class RD>
__device__ __forceinline__ Fr eval_lc(const ProgView& pv, const RD& rd, const Fr* __restrict__ W, uint32_t& p) {
    const uint32_t n = rd(p), ci = rd(p + 1);
    p += 2;
    Fr acc = ci != 0xffffffffu ? pv.consts_can[ci] : Fr::zero();
    if (n == 0) return acc;
    // one term of lookahead: the next wire is in flight while this one is folded in
    uint32_t w = rd(p), c = rd(p + 1);
    Fr v = W[w];
    for (uint32_t t = 0; t < n; t++) {
        const uint32_t c_cur = c;
        const Fr v_cur = v;
        p += 2;
        if (t + 1 < n) {
            w = rd(p);
            c = rd(p + 1);
            v = W[w];
        }
        lc_term(pv, acc, c_cur, v_cur);
    }
    return acc;
}

// the same LC evaluated by a whole warp: lane j folds terms j, j+32, ...; butterfly sum over the lanes.
// Every lane returns the value.
__device__ __forceinline__ Fr eval_lc_warp(const ProgView& pv, const Fr* __restrict__ W, uint32_t& p, uint32_t lane) {
    const uint32_t n = pv.code[p], ci = pv.code[p + 1];
    p += 2;
    Fr acc = (ci != 0xffffffffu && lane == 0) ? pv.consts_can[ci] : Fr::zero();
    // four terms per lane at a time: their code words leave together, then their wires -- two memory round trips
    // per chunk instead of eight
    for (uint32_t t0 = lane; t0 < n; t0 += 128) {
        uint32_t w[4], c[4];
        Fr v[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const uint32_t t = t0 + 32 * k;
            if (t < n) {
                w[k] = pv.code[p + 2 * t];
                c[k] = pv.code[p + 2 * t + 1];
            }
        }
#pragma unroll
        for (int k = 0; k < 4; k++)
            if (t0 + 32 * k < n) v[k] = W[w[k]];
#pragma unroll
        for (int k = 0; k < 4; k++)
            if (t0 + 32 * k < n) lc_term(pv, acc, c[k], v[k]);
    }
    p += 2 * n;
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        Fr o;
#pragma unroll
        for (int i = 0; i < 8; i++) o.v[i] = __shfl_xor_sync(0xffffffffu, acc.v[i], off);
        acc = acc + o;
    }
    return acc;
}

// a * b for canonical operands, canonical result
__device__ __forceinline__ Fr mul_canonical(const Fr& a, const Fr& b) {
    if (a.is_zero() || b.is_zero()) return Fr::zero();
    const bool sa = fits_u32(a), sb = fits_u32(b);
    if (sa && a.v[0] == 1) return b;
    if (sb && b.v[0] == 1) return a;
    if (sa && sb) {  // bytes, words, positions: the product fits 64 bits, far below r
        const uint64_t pr = (uint64_t)a.v[0] * b.v[0];
        Fr o = Fr::zero();
        o.v[0] = (uint32_t)pr;
        o.v[1] = (uint32_t)(pr >> 32);
        return o;
    }
    return wit_mul(wit_mul(a, b), Fr::r2());
}

__device__ __forceinline__ Fr inv_or_zero(const ProgView& pv, const Fr& v) {
    if (v.is_zero()) return v;
    if (fits_u32(v) && v.v[0] <= NZ_INV_TAB) return pv.invtab[v.v[0]];
    const Fr m = v.neg();
    if (fits_u32(m) && m.v[0] <= NZ_INV_TAB) return pv.invtab[m.v[0]].neg();  // 1/(-k) = -(1/k)
    return wit_inv(v);
}

// ---- word-level SHA-2 steps (builder.OP_SHAROUND / OP_SHASCHED; circuits/nzcptpl.circom:509-516 Sha256Var(3),
// :577-580 Sha512(512)) -----------------------------------------------------------------------------------------
// One warp per step.  The step's input words are gathered bit by bit (a bit is a wire or a short LC of wires) with
// warp ballots, the round / schedule step is computed on n-bit words, and every bit wire the circuit has for the step
// (Xor3's b*c products and outputs, Ch, Maj's product and output, the carry-extended BinSum decompositions) is
// written at its place: the same wires the generic instructions of the step write (the oracle VM runs those).
template <class WT>
struct ShaWord;
template <>
struct ShaWord<uint32_t> {
    static constexpr uint32_t N = 32;
    typedef uint64_t Wide;
};
template <>
struct ShaWord<uint64_t> {
    static constexpr uint32_t N = 64;
    typedef unsigned __int128 Wide;
};

template <class WT>
__device__ __forceinline__ WT sha_rotr(WT x, uint32_t r) {
    constexpr uint32_t N = ShaWord<WT>::N;
    r &= N - 1;
    return r ? (WT)((x >> r) | (x << (N - r))) : x;
}
__device__ __forceinline__ void sha_put(Fr* __restrict__ W, uint32_t idx, uint32_t bit) {
    Fr o = Fr::zero();
    o.v[0] = bit;
    W[idx] = o;
}
// bit `i` of input word `k` of the instruction at p: refs start at `refs`
__device__ __forceinline__ uint32_t sha_in_bit(const ProgView& pv, const Fr* __restrict__ W, uint32_t p, uint32_t refs,
                                               uint32_t n, uint32_t k, uint32_t i) {
    const uint32_t ref = pv.code[refs + k * n + i];
    if (!(ref & 0x80000000u)) return W[ref].v[0] & 1u;
    uint32_t q = p + (ref & 0x7fffffffu);
    const GlobalCode gc{pv.code};
    return eval_lc(pv, gc, W, q).v[0] & 1u;
}
template <class WT, int NW>
__device__ __forceinline__ void sha_gather(const ProgView& pv, const Fr* __restrict__ W, uint32_t p, uint32_t refs,
                                           uint32_t lane, WT* word) {
    constexpr uint32_t N = ShaWord<WT>::N;
    uint32_t lo[NW], hi[NW];
#pragma unroll
    for (int k = 0; k < NW; k++) {
        lo[k] = sha_in_bit(pv, W, p, refs, N, k, lane);
        hi[k] = N == 64 ? sha_in_bit(pv, W, p, refs, N, k, lane + 32) : 0u;
    }
#pragma unroll
    for (int k = 0; k < NW; k++) {
        const uint32_t l = __ballot_sync(0xffffffffu, lo[k]);
        const uint32_t h = N == 64 ? __ballot_sync(0xffffffffu, hi[k]) : 0u;
        word[k] = (WT)(((uint64_t)h << 32) | l);
    }
}
// Xor3 block: for i < n_mid (mid, out) interleaved, then out only
template <class WT>
__device__ __forceinline__ void sha_put_xor3(Fr* __restrict__ W, uint32_t base, WT mid, WT out, uint32_t n_mid, uint32_t lane) {
    constexpr uint32_t N = ShaWord<WT>::N;
    for (uint32_t i = lane; i < N; i += 32) {
        if (i < n_mid) {
            sha_put(W, base + 2 * i, (uint32_t)((mid >> i) & 1));
            sha_put(W, base + 2 * i + 1, (uint32_t)((out >> i) & 1));
        } else {
            sha_put(W, base + 2 * n_mid + (i - n_mid), (uint32_t)((out >> i) & 1));
        }
    }
}
template <class WIDE>
__device__ __forceinline__ void sha_put_bits(Fr* __restrict__ W, uint32_t base, WIDE v, uint32_t count, uint32_t lane) {
    for (uint32_t j = lane; j < count; j += 32) sha_put(W, base + j, (uint32_t)((v >> j) & 1));
}

// one round on words: writes the round's 11 N + 6 bit wires at w0, returns the new e and a through en / an
template <class WT>
__device__ __forceinline__ void sha_round_emit(Fr* __restrict__ W, uint32_t w0, WT A, WT B, WT C, WT D, WT E, WT F, WT G, WT H,
                                               WT Wt, WT K, uint32_t r1a, uint32_t r1b, uint32_t r1c, uint32_t r0a,
                                               uint32_t r0b, uint32_t r0c, uint32_t lane, WT& e_new, WT& a_new) {
    constexpr uint32_t N = ShaWord<WT>::N;
    typedef typename ShaWord<WT>::Wide Wide;
    const WT e2 = sha_rotr(E, r1b), e3 = sha_rotr(E, r1c);
    const WT S1 = sha_rotr(E, r1a) ^ e2 ^ e3;
    const WT ch = (E & F) ^ (~E & G);
    const Wide t1 = (Wide)H + S1 + ch + K + Wt;
    const WT a2 = sha_rotr(A, r0b), a3 = sha_rotr(A, r0c);
    const WT S0 = sha_rotr(A, r0a) ^ a2 ^ a3;
    const WT mj = (A & B) ^ (A & C) ^ (B & C);
    const Wide t2 = (Wide)S0 + mj;
    const Wide en = (Wide)D + (WT)t1;
    const Wide an = (Wide)(WT)t1 + (WT)t2;
    sha_put_xor3<WT>(W, w0, e2 & e3, S1, N, lane);
    sha_put_bits<WT>(W, w0 + 2 * N, ch, N, lane);
    sha_put_bits<Wide>(W, w0 + 3 * N, t1, N + 3, lane);
    sha_put_xor3<WT>(W, w0 + 4 * N + 3, a2 & a3, S0, N, lane);
    sha_put_xor3<WT>(W, w0 + 6 * N + 3, B & C, mj, N, lane);
    sha_put_bits<Wide>(W, w0 + 8 * N + 3, t2, N + 1, lane);
    sha_put_bits<Wide>(W, w0 + 9 * N + 4, en, N + 1, lane);
    sha_put_bits<Wide>(W, w0 + 10 * N + 5, an, N + 1, lane);
    e_new = (WT)en;
    a_new = (WT)an;
}
// one message-schedule step on words: writes its 5 N + 2 - r1c - r0c bit wires at w0, returns w[t]
template <class WT>
__device__ __forceinline__ WT sha_sched_emit(Fr* __restrict__ W, uint32_t w0, WT x2, WT x7, WT x15, WT x16, uint32_t r1a,
                                             uint32_t r1b, uint32_t r1c, uint32_t r0a, uint32_t r0b, uint32_t r0c,
                                             uint32_t lane) {
    constexpr uint32_t N = ShaWord<WT>::N;
    typedef typename ShaWord<WT>::Wide Wide;
    const WT b1 = sha_rotr(x2, r1b), c1 = (WT)(x2 >> r1c);
    const WT s1 = sha_rotr(x2, r1a) ^ b1 ^ c1;
    const WT b0 = sha_rotr(x15, r0b), c0 = (WT)(x15 >> r0c);
    const WT s0 = sha_rotr(x15, r0a) ^ b0 ^ c0;
    const Wide sum = (Wide)s1 + x7 + s0 + x16;
    const uint32_t o_s0 = 2 * N - r1c, o_w = o_s0 + 2 * N - r0c;
    sha_put_xor3<WT>(W, w0, b1 & c1, s1, N - r1c, lane);
    sha_put_xor3<WT>(W, w0 + o_s0, b0 & c0, s0, N - r0c, lane);
    sha_put_bits<Wide>(W, w0 + o_w, sum, N + 2, lane);
    return (WT)sum;
}

// out of line: the interpreter's own loop is latency bound and must keep its registers.  TAG gives every kernel its
// own copy (ptxas 12.9 segfaults when two kernels of this translation unit share an out-of-line routine).
template <class WT, uint32_t TAG>
__device__ __noinline__ void sha_step_warp(const ProgView& pv, Fr* __restrict__ W, uint32_t p, uint32_t lane, uint32_t op) {
    const uint32_t r1a = pv.code[p + 2], r1b = pv.code[p + 3], r1c = pv.code[p + 4];
    const uint32_t r0a = pv.code[p + 5], r0b = pv.code[p + 6], r0c = pv.code[p + 7];
    if (op == OP_SHAROUND) {
        const WT K = (WT)(((uint64_t)pv.code[p + 9] << 32) | pv.code[p + 8]);
        const uint32_t w0 = pv.code[p + 10], refs = p + 12;
        WT x[9], en, an;  // a b c d e f g h w
        sha_gather<WT, 9>(pv, W, p, refs, lane, x);
        sha_round_emit<WT>(W, w0, x[0], x[1], x[2], x[3], x[4], x[5], x[6], x[7], x[8], K, r1a, r1b, r1c, r0a, r0b, r0c, lane, en, an);
    } else {
        const uint32_t w0 = pv.code[p + 8], refs = p + 10;
        WT x[4];  // w[t-2] w[t-7] w[t-15] w[t-16]
        sha_gather<WT, 4>(pv, W, p, refs, lane, x);
        sha_sched_emit<WT>(W, w0, x[0], x[1], x[2], x[3], r1a, r1b, r1c, r0a, r0b, r0c, lane);
    }
}

// A whole SHA-256 compression from round r_start on (builder.OP_SHABLOCK): the eight state words and the sixteen
// message words are gathered once, the state lives in registers across the rounds, the 16-word schedule window is
// spread over the lanes (lane t & 15 holds w[t]) and read with shuffles.  One warp, one level.
template <uint32_t TAG>
__device__ __noinline__ void sha_block_warp(const ProgView& pv, Fr* __restrict__ W, uint32_t p, uint32_t lane) {
    const uint32_t* cd = pv.code + p;
    const uint32_t r_start = cd[14], rounds = cd[15];
    const uint32_t sched = p + 16, rnd = sched + (rounds - 16), kk = rnd + (rounds - r_start), refs = kk + (rounds - r_start);
    uint32_t st[8], win = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) st[k] = __ballot_sync(0xffffffffu, sha_in_bit(pv, W, p, refs, 32, k, lane));
    for (uint32_t k = 0; k < 16; k++) {  // message word k -> lane k of the window
        const uint32_t v = __ballot_sync(0xffffffffu, sha_in_bit(pv, W, p, refs, 32, 8 + k, lane));
        if (lane == k) win = v;
    }
    uint32_t a = st[0], b = st[1], c = st[2], d = st[3], e = st[4], f = st[5], g = st[6], h = st[7];
    // r_start <= 16 (loader): every schedule step t >= 16 is computed right before its round
    for (uint32_t t = r_start; t < rounds; t++) {
        if (t >= 16) {
            const uint32_t x2 = __shfl_sync(0xffffffffu, win, (t - 2) & 15), x7 = __shfl_sync(0xffffffffu, win, (t - 7) & 15);
            const uint32_t x15 = __shfl_sync(0xffffffffu, win, (t - 15) & 15), x16 = __shfl_sync(0xffffffffu, win, t & 15);
            const uint32_t wt = sha_sched_emit<uint32_t>(W, pv.code[sched + (t - 16)], x2, x7, x15, x16, cd[8], cd[9], cd[10], cd[11],
                                                         cd[12], cd[13], lane);
            if (lane == (t & 15)) win = wt;
        }
        const uint32_t wt = __shfl_sync(0xffffffffu, win, t & 15);
        uint32_t en, an;
        sha_round_emit<uint32_t>(W, pv.code[rnd + (t - r_start)], a, b, c, d, e, f, g, h, wt, pv.code[kk + (t - r_start)], cd[2], cd[3],
                                 cd[4], cd[5], cd[6], cd[7], lane, en, an);
        h = g; g = f; f = e; e = en;
        d = c; c = b; b = a; a = an;
    }
}

// The rounds of a compression from round r_start on (builder.OP_SHAROUNDS), for compressions whose message schedule
// keeps its own instructions: the state is gathered once and lives in registers, w[t] is gathered round by round.
template <class WT, uint32_t TAG>
__device__ __noinline__ void sha_rounds_warp(const ProgView& pv, Fr* __restrict__ W, uint32_t p, uint32_t lane) {
    constexpr uint32_t N = ShaWord<WT>::N;
    const uint32_t* cd = pv.code + p;
    const uint32_t r_start = cd[8], rounds = cd[9], nr = rounds - r_start;
    const uint32_t rnd = p + 10, kk = rnd + nr, refs = kk + 2 * nr;
    WT st[8];
    sha_gather<WT, 8>(pv, W, p, refs, lane, st);
    WT a = st[0], b = st[1], c = st[2], d = st[3], e = st[4], f = st[5], g = st[6], h = st[7];
    for (uint32_t i = 0; i < nr; i++) {
        WT wt;
        sha_gather<WT, 1>(pv, W, p, refs + (8 + i) * N, lane, &wt);
        const WT K = (WT)(((uint64_t)pv.code[kk + 2 * i + 1] << 32) | pv.code[kk + 2 * i]);
        WT en, an;
        sha_round_emit<WT>(W, pv.code[rnd + i], a, b, c, d, e, f, g, h, wt, K, cd[2], cd[3], cd[4], cd[5], cd[6], cd[7], lane, en, an);
        h = g; g = f; f = e; e = en;
        d = c; c = b; b = a; a = an;
    }
}

// ---- QuinSelector(N) as one instruction (builder.OP_QUINSEL; circuits/quinSelector.circom:26-41, the GetV of
// cbortpl.circom:79-90 and every selector of the CBOR walk).  For each choice i the circuit has three signals:
// eqs[i].inv (IsZero's hint: 1/(i - index) or 0), eqs[i].out (i == index) and sums[i] (the running sum, which is
// in[index] from i = index on and 0 before).  One warp writes all 3 N of them from the index and ONE input value.
template <uint32_t TAG>
__device__ __noinline__ void quinsel_warp(const ProgView& pv, Fr* __restrict__ W, uint32_t p, uint32_t lane) {
    const uint32_t N = pv.code[p + 1], dsts = p + 2, refs = dsts + 2 * N;
    const GlobalCode gc{pv.code};
    uint32_t q = refs + N;
    const Fr index = eval_lc(pv, gc, W, q);  // every lane: same addresses, one transaction
    const bool valid = fits_u32(index) && index.v[0] < N;
    Fr val = Fr::zero();
    if (valid) {
        const uint32_t ref = pv.code[refs + index.v[0]];
        if (!(ref & 0x80000000u)) {
            val = W[ref];
        } else {
            uint32_t r = p + (ref & 0x7fffffffu);
            val = eval_lc(pv, gc, W, r);
        }
    }
    for (uint32_t i = lane; i < N; i += 32) {
        const uint32_t eq_w = pv.code[dsts + 2 * i], sum_w = pv.code[dsts + 2 * i + 1];
        Fr d = Fr::zero();
        d.v[0] = i;
        d = d - index;
        Fr eq = Fr::zero();
        eq.v[0] = d.is_zero() ? 1u : 0u;
        W[eq_w - 1] = inv_or_zero(pv, d);
        W[eq_w] = eq;
        if (sum_w != 0xffffffffu) W[sum_w] = (valid && i >= index.v[0]) ? val : Fr::zero();
    }
}

// Executes the instruction whose words rd() serves from offset p.  WARP: all 32 lanes cooperate on its LCs (lane 0
// commits; code stream only); else one thread.
template <bool WARP, class RD, uint32_t TAG = 0>
__device__ __forceinline__ bool exec_instr(const ProgView& pv, const RD& rd, Fr* __restrict__ W, uint32_t p, uint32_t lane) {
    const uint32_t op = rd(p) & 0xffu;
    auto LC = [&](uint32_t& q) { return WARP ? eval_lc_warp(pv, W, q, lane) : eval_lc(pv, rd, W, q); };
    const bool commit = !WARP || lane == 0;
    if (op == OP_LIN) {
        const uint32_t dst = rd(p + 1);
        p += 2;
        const Fr v = LC(p);
        if (commit) W[dst] = v;
    } else if (op == OP_MUL) {
        const uint32_t dst = rd(p + 1);
        p += 2;
        const Fr a = LC(p);
        const Fr b = LC(p);
        const Fr c = LC(p);
        if (commit) W[dst] = mul_canonical(a, b) + c;
    } else if (op == OP_BITS) {
        const uint32_t dst = rd(p + 1), src = rd(p + 2), n = rd(p + 3);
        const Fr v = W[src];
        for (uint32_t k = WARP ? lane : 0; k < n; k += WARP ? 32 : 1) {
            Fr bit = Fr::zero();
            bit.v[0] = k < 256 ? (v.v[k >> 5] >> (k & 31)) & 1u : 0u;
            W[dst + k] = bit;
        }
    } else if (op == OP_BITSLC) {  // bits of the value of an LC: the sum and its decomposition in one instruction
        const uint32_t dst = rd(p + 1), n = rd(p + 2);
        p += 3;
        const Fr v = LC(p);  // warp mode: every lane holds the value
        for (uint32_t k = WARP ? lane : 0; k < n; k += WARP ? 32 : 1) {
            Fr bit = Fr::zero();
            bit.v[0] = k < 256 ? (v.v[k >> 5] >> (k & 31)) & 1u : 0u;
            W[dst + k] = bit;
        }
    } else if (op == OP_INV) {
        if (commit) W[rd(p + 1)] = inv_or_zero(pv, W[rd(p + 2)]);
    } else if (op == OP_QUINSEL) {  // always scheduled as a one-warp instruction (loader)
        if (WARP) quinsel_warp<TAG>(pv, W, p, lane);
    } else if (op == OP_SHABLOCK) {
        if (WARP) sha_block_warp<TAG>(pv, W, p, lane);
    } else if (op == OP_SHAROUNDS) {
        if (WARP) {
            if (pv.code[p + 1] == 64) sha_rounds_warp<uint64_t, TAG>(pv, W, p, lane);
            else sha_rounds_warp<uint32_t, TAG>(pv, W, p, lane);
        }
    } else if (op == OP_SHAROUND || op == OP_SHASCHED) {
        if (WARP) {
            if (pv.code[p + 1] == 64) sha_step_warp<uint64_t, TAG>(pv, W, p, lane, op);
            else sha_step_warp<uint32_t, TAG>(pv, W, p, lane, op);
        }
    } else {  // OP_ASSERT
        p += 1;
        const Fr a = LC(p);
        const Fr b = LC(p);
        const Fr c = LC(p);
        if (commit && mul_canonical(a, b) != c) return true;
    }
    return false;
}

// LIN / MUL / ASSERT whose whole encoding sits in this thread's shared-memory record: every wire the instruction
// reads is requested before the first one is used, so the three LCs of a product cost one memory round trip, not
// three.  (BITS / INV and anything longer go through exec_instr.)
template <class RD>
__device__ __forceinline__ bool exec_short(const ProgView& pv, const RD& rd, Fr* __restrict__ W, uint32_t op) {
    const uint32_t base_a = (op == OP_ASSERT ? 1u : 2u) + 2u;  // first term word of LC A
    const uint32_t na = rd(base_a - 2), ci_a = rd(base_a - 1);
    uint32_t nb = 0, nc = 0, ci_b = 0xffffffffu, ci_c = 0xffffffffu;
    if (op != OP_LIN) {
        const uint32_t hb = base_a + 2 * na;
        nb = rd(hb);
        ci_b = rd(hb + 1);
        const uint32_t hc = hb + 2 + 2 * nb;
        nc = rd(hc);
        ci_c = rd(hc + 1);
    }
    const uint32_t nab = na + nb, nt = nab + nc;
    auto pos = [&](uint32_t t) { return base_a + 2 * t + (t >= na ? 2u : 0u) + (t >= nab ? 2u : 0u); };
    constexpr int BATCH = 4;
    Fr v[BATCH];
#pragma unroll
    for (int k = 0; k < BATCH; k++)
        if ((uint32_t)k < nt) v[k] = W[rd(pos(k))];
    Fr A = ci_a != 0xffffffffu ? pv.consts_can[ci_a] : Fr::zero();
    Fr Bv = ci_b != 0xffffffffu ? pv.consts_can[ci_b] : Fr::zero();
    Fr Cv = ci_c != 0xffffffffu ? pv.consts_can[ci_c] : Fr::zero();
    auto fold = [&](uint32_t t, const Fr& x) {
        const uint32_t c = rd(pos(t) + 1);
        if (t < na) lc_term(pv, A, c, x);
        else if (t < nab) lc_term(pv, Bv, c, x);
        else lc_term(pv, Cv, c, x);
    };
#pragma unroll
    for (int k = 0; k < BATCH; k++)
        if ((uint32_t)k < nt) fold(k, v[k]);
    for (uint32_t t = BATCH; t < nt; t++) fold(t, W[rd(pos(t))]);
    if (op == OP_LIN) {
        W[rd(1)] = A;
        return false;
    }
    const Fr ab = mul_canonical(A, Bv);
    if (op == OP_MUL) {
        W[rd(1)] = ab + Cv;
        return false;
    }
    return ab != Cv;  // OP_ASSERT
}

// One CTA per pass.  Within a level the host has put the "long" instructions (many LC terms, or wide bit
// decompositions) first: those run one per warp from the code stream.  The rest run one per thread; the encoding of
// the instruction a thread will execute in the NEXT level is copied asynchronously (cp.async) from the flat record
// array into its shared-memory slot while the current level runs, so a level's critical path is wire loads,
// arithmetic, store, barrier -- no instruction fetch.
template <uint32_t T, uint32_t CTAS_PER_SM>
__global__ void __launch_bounds__(T, CTAS_PER_SM) k_witness(ProgView pv, const Fr* __restrict__ inputs, Fr* __restrict__ wires,
                                                            int32_t* __restrict__ status, uint32_t B) {
    extern __shared__ uint4 s_rec[];  // [2][8][T]
    constexpr uint32_t WIT_THREADS = T;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = WIT_THREADS / 32;
    const GlobalCode gcode{pv.code};
    auto slot = [&](uint32_t buf, uint32_t k8) { return s_rec + ((size_t)buf * 8 + k8) * WIT_THREADS + threadIdx.x; };
    auto prefetch = [&](uint32_t buf, uint32_t i) {
        const uint4* src = pv.rec + (size_t)i * 8;
#pragma unroll
        for (uint32_t k8 = 0; k8 < 8; k8++) __pipeline_memcpy_async(slot(buf, k8), src + k8, 16);
    };
    for (uint32_t pass = blockIdx.x; pass < B; pass += gridDim.x) {
        Fr* W = wires + (size_t)pass * pv.n_total;
        const Fr* in = inputs + (size_t)pass * pv.n_in;
        // circom_runtime stores Fr.e(value): any 256-bit input is taken mod r (the multiply needs reduced operands)
        for (uint32_t i = threadIdx.x; i < pv.n_in; i += blockDim.x) W[1 + pv.n_out + i] = fr_reduce_256(in[i]);
        if (threadIdx.x == 0) {
            Fr one = Fr::zero();
            one.v[0] = 1;
            W[0] = one;
        }
        bool failed = false;
        uint32_t lo = pv.lstart[0], hi = pv.n_levels ? pv.lstart[1] : 0, nl = pv.n_levels ? pv.nlong[0] : 0;
        if (lo + nl + threadIdx.x < hi) prefetch(0, lo + nl + threadIdx.x);
        __pipeline_commit();
        __syncthreads();
        for (uint32_t l = 0; l < pv.n_levels; l++) {
            const uint32_t buf = l & 1;
            // next level's bounds, and this thread's instruction of it, are fetched while this level executes
            uint32_t hi_n = 0, nl_n = 0;
            if (l + 1 < pv.n_levels) {
                hi_n = pv.lstart[l + 2];
                nl_n = pv.nlong[l + 1];
            }
            if (hi + nl_n + threadIdx.x < hi_n) prefetch(buf ^ 1, hi + nl_n + threadIdx.x);
            __pipeline_commit();
            for (uint32_t i = lo + warp; i < lo + nl; i += n_warps) failed |= exec_instr<true, GlobalCode, T * 16 + CTAS_PER_SM>(pv, gcode, W, pv.ioff[i], lane);
            uint32_t i = lo + nl + threadIdx.x;
            if (i < hi) {
                __pipeline_wait_prior(1);  // everything but the copy just issued has landed: this level's record
                const SmemCode<T> scode{reinterpret_cast<const uint32_t*>(slot(buf, 0))};
                const uint32_t w0 = scode(0), op = w0 & 0xffu;
                if (w0 & REC_LONG) failed |= exec_instr<false>(pv, gcode, W, scode(1), lane);
                else if (op == OP_BITS || op == OP_INV || op == OP_BITSLC) failed |= exec_instr<false>(pv, scode, W, 0, lane);
                else failed |= exec_short(pv, scode, W, op);
                for (i += WIT_THREADS; i < hi; i += WIT_THREADS) failed |= exec_instr<false>(pv, gcode, W, pv.ioff[i], lane);
            }
            __syncthreads();
            lo = hi;
            hi = hi_n;
            nl = nl_n;
        }
        __pipeline_wait_prior(0);
        if (failed) atomicExch(&status[pass], NZCB_E_ASSERT);
        __syncthreads();
    }
}

__global__ void k_invtab(Fr* tab) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k > NZ_INV_TAB) return;
    tab[k] = k == 0 ? Fr::zero() : Fr::from_u64(k).inv().from_mont();
}

__global__ void k_consts_to_mont(Fr* c, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) c[i] = c[i].to_mont();
}

}  // namespace

extern "C" void nzcb_circuit_free(nzcb_circuit* c) {
    if (!c) return;
    if (c->ctx) {
        cudaSetDevice(c->ctx->device);
        cudaStreamSynchronize(c->ctx->stream);
    }
    cudaFree(c->d_consts);
    cudaFree(c->d_consts_can);
    cudaFree(c->d_nlong);
    cudaFree(c->d_rec);
    cudaFree(c->d_ioff);
    cudaFree(c->d_lstart);
    cudaFree(c->d_code);
    cudaFree(c->d_invtab);
    delete c;
}

extern "C" int32_t nzcb_circuit_info(const nzcb_circuit* c, uint32_t* n_witness, uint32_t* n_inputs, uint32_t* n_outputs) {
    if (!c) return NZCB_E_INVALID;
    if (n_witness) *n_witness = c->n_witness;
    if (n_inputs) *n_inputs = c->n_in;
    if (n_outputs) *n_outputs = c->n_out;
    return 0;
}

#define WC_CUDA(call)                                                                                      \
    do {                                                                                                   \
        cudaError_t e__ = (call);                                                                          \
        if (e__ != cudaSuccess) {                                                                          \
            ctx->fail(NZCB_E_CUDA, "CUDA error %s at %s:%d", cudaGetErrorString(e__), __FILE__, __LINE__); \
            nzcb_circuit_free(c);                                                                          \
            return NZCB_E_CUDA;                                                                            \
        }                                                                                                  \
    } while (0)

extern "C" int32_t nzcb_circuit_load(nzcb_ctx* ctx, const uint8_t* data, size_t len, nzcb_circuit** out) {
    if (!ctx || !data || !out) return NZCB_E_INVALID;
    *out = nullptr;
    if (len < 40 || memcmp(data, "NZWP", 4) != 0) return ctx->fail(NZCB_E_INVALID, "witness program: bad magic");
    uint32_t h[9];
    memcpy(h, data + 4, 36);
    if (h[0] != 1) return ctx->fail(NZCB_E_INVALID, "witness program: unsupported version %u", h[0]);
    nzcb_circuit* c = new nzcb_circuit();
    c->ctx = ctx;
    c->n_total = h[1]; c->n_witness = h[2]; c->n_out = h[3]; c->n_in = h[4];
    c->n_consts = h[5]; c->n_instr = h[6]; c->n_levels = h[7]; c->n_code = h[8];
    const size_t need = 40 + (size_t)c->n_consts * 32 + (size_t)c->n_instr * 4 + ((size_t)c->n_levels + 1) * 4 +
                        (size_t)c->n_code * 4;
    if (need != len || c->n_witness > c->n_total || 1 + (uint64_t)c->n_out + c->n_in > c->n_witness || c->n_consts < 2) {
        delete c;
        return ctx->fail(NZCB_E_INVALID, "witness program: inconsistent header");
    }
    const uint8_t* p_consts = data + 40;
    const uint8_t* p_ioff = p_consts + (size_t)c->n_consts * 32;
    const uint8_t* p_lstart = p_ioff + (size_t)c->n_instr * 4;
    const uint8_t* p_code = p_lstart + ((size_t)c->n_levels + 1) * 4;
    // validate once on the host so the kernel can trust every index; measure every instruction
    std::vector<uint32_t> ioff_sorted(c->n_instr), nlong(std::max<uint32_t>(1, c->n_levels), 0), recs;
    {
        std::vector<uint32_t> code(c->n_code), ioff(c->n_instr), ls(c->n_levels + 1), weight(c->n_instr, 0);
        memcpy(code.data(), p_code, (size_t)c->n_code * 4);
        memcpy(ioff.data(), p_ioff, (size_t)c->n_instr * 4);
        memcpy(ls.data(), p_lstart, ((size_t)c->n_levels + 1) * 4);
        bool ok = ls[0] == 0 && ls[c->n_levels] == c->n_instr;
        for (uint32_t l = 0; ok && l < c->n_levels; l++) ok = ls[l] <= ls[l + 1];
        uint32_t cur_weight = 0;  // longest LC of the instruction being checked (or its bit count)
        auto lc_ok = [&](uint32_t& p) {
            if (p + 2 > c->n_code) return false;
            const uint32_t n = code[p], ci = code[p + 1];
            if (ci != 0xffffffffu && ci >= c->n_consts) return false;
            p += 2;
            if ((uint64_t)p + 2ull * n > c->n_code) return false;
            for (uint32_t t = 0; t < n; t++, p += 2)
                if (code[p] >= c->n_total || code[p + 1] >= c->n_consts) return false;
            cur_weight = std::max(cur_weight, n);
            return true;
        };
        for (uint32_t i = 0; ok && i < c->n_instr; i++) {
            uint32_t p = ioff[i];
            if (p + 1 > c->n_code) { ok = false; break; }
            const uint32_t op = code[p];
            cur_weight = 0;
            if (op == OP_LIN) {
                ok = p + 2 <= c->n_code && code[p + 1] < c->n_total; p += 2; ok = ok && lc_ok(p);
            } else if (op == OP_MUL) {
                ok = p + 2 <= c->n_code && code[p + 1] < c->n_total; p += 2; ok = ok && lc_ok(p) && lc_ok(p) && lc_ok(p);
            } else if (op == OP_BITS) {
                ok = p + 4 <= c->n_code && code[p + 2] < c->n_total && (uint64_t)code[p + 1] + code[p + 3] <= c->n_total;
                if (ok) cur_weight = code[p + 3];
            } else if (op == OP_BITSLC) {
                ok = p + 3 <= c->n_code && (uint64_t)code[p + 1] + code[p + 2] <= c->n_total;
                const uint32_t n_bits = ok ? code[p + 2] : 0;
                p += 3;
                ok = ok && lc_ok(p);
                if (ok) cur_weight = std::max(cur_weight, n_bits);
            } else if (op == OP_INV) {
                ok = p + 3 <= c->n_code && code[p + 1] < c->n_total && code[p + 2] < c->n_total;
            } else if (op == OP_ASSERT) {
                p += 1; ok = lc_ok(p) && lc_ok(p) && lc_ok(p);
            } else if (op == OP_SHAROUNDS) {
                ok = (uint64_t)p + 10 <= c->n_code && (code[p + 1] == 32 || code[p + 1] == 64);
                const uint32_t n = ok ? code[p + 1] : 0, r_start = ok ? code[p + 8] : 0, rounds = ok ? code[p + 9] : 0;
                ok = ok && rounds <= 128 && r_start < rounds;
                if (ok) {
                    const uint32_t nr = rounds - r_start, rnd = p + 10, kk = rnd + nr, refs = kk + 2 * nr, n_refs = (8 + nr) * n;
                    for (uint32_t k = 2; k < 8; k++) ok = ok && code[p + k] < n;
                    ok = ok && (uint64_t)refs + n_refs <= c->n_code;
                    for (uint32_t k = 0; ok && k < nr; k++) ok = (uint64_t)code[rnd + k] + 11 * n + 6 <= c->n_total;
                    for (uint32_t k = 0; ok && k < n_refs; k++) {
                        const uint32_t ref = code[refs + k];
                        if (ref & 0x80000000u) {
                            uint32_t q = p + (ref & 0x7fffffffu);
                            ok = q >= refs + n_refs && lc_ok(q);
                        } else {
                            ok = ref < c->n_total;
                        }
                    }
                }
                cur_weight = 1u << 20;  // one warp, first in its level
            } else if (op == OP_SHABLOCK) {
                ok = (uint64_t)p + 16 <= c->n_code && code[p + 1] == 32;
                const uint32_t r_start = ok ? code[p + 14] : 0, rounds = ok ? code[p + 15] : 0;
                ok = ok && rounds >= 17 && rounds <= 128 && r_start <= 16;
                if (ok) {
                    for (uint32_t k = 2; k < 14; k++) ok = ok && code[p + k] < 32;
                    const uint32_t sched = p + 16, rnd = sched + (rounds - 16), kk = rnd + (rounds - r_start);
                    const uint32_t refs = kk + (rounds - r_start);
                    ok = ok && (uint64_t)refs + 24 * 32 <= c->n_code;
                    const uint32_t s_size = ok ? 5 * 32 + 2 - code[p + 10] - code[p + 13] : 0;
                    for (uint32_t k = 0; ok && k < rounds - 16; k++) ok = (uint64_t)code[sched + k] + s_size <= c->n_total;
                    for (uint32_t k = 0; ok && k < rounds - r_start; k++) ok = (uint64_t)code[rnd + k] + 11 * 32 + 6 <= c->n_total;
                    for (uint32_t k = 0; ok && k < 24 * 32; k++) {
                        const uint32_t ref = code[refs + k];
                        if (ref & 0x80000000u) {
                            uint32_t q = p + (ref & 0x7fffffffu);
                            ok = q >= refs + 24 * 32 && lc_ok(q);
                        } else {
                            ok = ref < c->n_total;
                        }
                    }
                }
                cur_weight = 1u << 20;  // one warp, first in its level
            } else if (op == OP_QUINSEL) {
                ok = (uint64_t)p + 2 <= c->n_code;
                const uint32_t N = ok ? code[p + 1] : 0;
                ok = ok && N >= 1 && N <= (1u << 20) && (uint64_t)p + 2 + 3ull * N <= c->n_code;
                if (ok) {
                    const uint32_t dsts = p + 2, refs = dsts + 2 * N;
                    for (uint32_t k = 0; ok && k < N; k++) {
                        const uint32_t ew = code[dsts + 2 * k], sw = code[dsts + 2 * k + 1];
                        ok = ew >= 1 && ew < c->n_total && (sw == 0xffffffffu || sw < c->n_total);
                    }
                    uint32_t q = refs + N;
                    ok = ok && lc_ok(q);
                    for (uint32_t k = 0; ok && k < N; k++) {
                        const uint32_t ref = code[refs + k];
                        if (ref & 0x80000000u) {
                            uint32_t r = p + (ref & 0x7fffffffu);
                            ok = r >= refs + N && lc_ok(r);
                        } else {
                            ok = ref < c->n_total;
                        }
                    }
                }
                cur_weight = 1u << 20;  // one warp, first in its level
            } else if (op == OP_SHAROUND || op == OP_SHASCHED) {
                const uint32_t hdr = op == OP_SHAROUND ? 12u : 10u, nw = op == OP_SHAROUND ? 9u : 4u;
                ok = (uint64_t)p + hdr <= c->n_code;
                const uint32_t n = ok ? code[p + 1] : 0;
                ok = ok && (n == 32 || n == 64) && (uint64_t)p + hdr + (uint64_t)nw * n <= c->n_code;
                if (ok) {
                    for (uint32_t k = 2; k < 8; k++) ok = ok && code[p + k] < n;
                    const uint32_t w0 = code[p + hdr - 2], size = code[p + hdr - 1];
                    const uint32_t want = op == OP_SHAROUND ? 11 * n + 6 : 5 * n + 2 - code[p + 4] - code[p + 7];
                    ok = ok && size == want && (uint64_t)w0 + size <= c->n_total;
                    for (uint32_t k = 0; ok && k < nw * n; k++) {
                        const uint32_t ref = code[p + hdr + k];
                        if (ref & 0x80000000u) {
                            uint32_t q = p + (ref & 0x7fffffffu);
                            ok = q >= p + hdr + nw * n && lc_ok(q);
                        } else {
                            ok = ref < c->n_total;
                        }
                    }
                }
                cur_weight = 1u << 20;  // one warp, first in its level
            } else ok = false;
            weight[i] = cur_weight;
        }
        if (!ok) {
            delete c;
            return ctx->fail(NZCB_E_INVALID, "witness program: malformed instruction stream");
        }
        // inside each level: long instructions (one warp each) first, by decreasing length
        std::vector<uint32_t> order;
        for (uint32_t l = 0; l < c->n_levels; l++) {
            order.clear();
            for (uint32_t i = ls[l]; i < ls[l + 1]; i++) order.push_back(i);
            std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return weight[a] > weight[b]; });
            uint32_t nl = 0;
            for (size_t k = 0; k < order.size(); k++) {
                ioff_sorted[ls[l] + k] = ioff[order[k]];
                if (weight[order[k]] > NZ_LONG_LC) nl++;
            }
            nlong[l] = nl;
        }
        // flat records, level order: an instruction's own words if they fit, else a pointer into the code stream
        recs.assign((size_t)std::max<uint32_t>(1, c->n_instr) * REC_WORDS, 0);
        for (uint32_t k = 0; k < c->n_instr; k++) {
            uint32_t* r = &recs[(size_t)k * REC_WORDS];
            const uint32_t p = ioff_sorted[k];
            const uint32_t op = code[p];
            uint32_t len;
            if (op == OP_SHAROUND || op == OP_SHASCHED || op == OP_QUINSEL || op == OP_SHABLOCK || op == OP_SHAROUNDS) len = REC_WORDS + 1;  // never a record: code stream
            else if (op == OP_BITS) len = 4;
            else if (op == OP_INV) len = 3;
            else if (op == OP_BITSLC) len = 3 + 2 + 2 * code[p + 3];
            else {
                uint32_t q = p + (op == OP_ASSERT ? 1 : 2);
                for (uint32_t j = 0; j < (op == OP_LIN ? 1u : 3u); j++) q += 2 + 2 * code[q];
                len = q - p;
            }
            if (len <= REC_WORDS) {
                for (uint32_t j = 0; j < len; j++) r[j] = code[p + j];
            } else {
                r[0] = op | REC_LONG;
                r[1] = p;
            }
        }
    }
    WC_CUDA(cudaSetDevice(ctx->device));
    WC_CUDA(cudaMalloc(&c->d_consts, (size_t)c->n_consts * 32));
    WC_CUDA(cudaMalloc(&c->d_consts_can, (size_t)c->n_consts * 32));
    WC_CUDA(cudaMalloc(&c->d_nlong, nlong.size() * 4));
    WC_CUDA(cudaMalloc(&c->d_rec, recs.size() * 4));
    WC_CUDA(cudaMemcpyAsync(c->d_rec, recs.data(), recs.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    WC_CUDA(cudaMemcpyAsync(c->d_consts_can, p_consts, (size_t)c->n_consts * 32, cudaMemcpyHostToDevice, ctx->stream));
    WC_CUDA(cudaMemcpyAsync(c->d_nlong, nlong.data(), nlong.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    WC_CUDA(cudaMalloc(&c->d_ioff, std::max<size_t>(4, (size_t)c->n_instr * 4)));
    WC_CUDA(cudaMalloc(&c->d_lstart, ((size_t)c->n_levels + 1) * 4));
    WC_CUDA(cudaMalloc(&c->d_code, std::max<size_t>(4, (size_t)c->n_code * 4)));
    WC_CUDA(cudaMalloc(&c->d_invtab, (NZ_INV_TAB + 1) * sizeof(Fr)));
    WC_CUDA(cudaMemcpyAsync(c->d_consts, p_consts, (size_t)c->n_consts * 32, cudaMemcpyHostToDevice, ctx->stream));
    WC_CUDA(cudaMemcpyAsync(c->d_ioff, ioff_sorted.data(), (size_t)c->n_instr * 4, cudaMemcpyHostToDevice, ctx->stream));
    WC_CUDA(cudaMemcpyAsync(c->d_lstart, p_lstart, ((size_t)c->n_levels + 1) * 4, cudaMemcpyHostToDevice, ctx->stream));
    WC_CUDA(cudaMemcpyAsync(c->d_code, p_code, (size_t)c->n_code * 4, cudaMemcpyHostToDevice, ctx->stream));
    k_consts_to_mont<<<div_up(c->n_consts, 256), 256, 0, ctx->stream>>>(c->d_consts, c->n_consts);
    k_invtab<<<div_up(NZ_INV_TAB + 1, 128), 128, 0, ctx->stream>>>(c->d_invtab);
    ctx->launches += 2;
    WC_CUDA(cudaGetLastError());
    WC_CUDA(cudaStreamSynchronize(ctx->stream));
    *out = c;
    return 0;
}

namespace nzcb {
// runs B passes; wires for pass i start at *d_wires + i * n_total (canonical LE).  Asynchronous on ctx->stream.
int witness_dev(nzcb_ctx* ctx, const nzcb_circuit* c, const Fr* d_inputs, size_t B, Fr* d_wires, int32_t* d_status) {
    ProgView pv;
    pv.consts = c->d_consts; pv.consts_can = c->d_consts_can; pv.ioff = c->d_ioff; pv.lstart = c->d_lstart;
    pv.nlong = c->d_nlong; pv.code = c->d_code; pv.invtab = c->d_invtab; pv.rec = c->d_rec;
    pv.n_total = c->n_total; pv.n_out = c->n_out; pv.n_in = c->n_in; pv.n_levels = c->n_levels;
    NZ_CUDA(ctx, cudaMemsetAsync(d_status, 0, B * sizeof(int32_t), ctx->stream));
    constexpr size_t smem = (size_t)2 * 8 * WIT_THREADS * sizeof(uint4);
    constexpr size_t smem_b = (size_t)2 * 8 * WIT_THREADS_BATCH * sizeof(uint4);
    constexpr size_t smem_64 = (size_t)2 * 8 * 64 * sizeof(uint4);
    static const bool attr_set = [] {
        return cudaFuncSetAttribute(k_witness<WIT_THREADS, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess &&
               cudaFuncSetAttribute(k_witness<WIT_THREADS_BATCH, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_b) == cudaSuccess &&
               cudaFuncSetAttribute(k_witness<WIT_THREADS_BATCH, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_b) == cudaSuccess &&
               cudaFuncSetAttribute(k_witness<64, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_64) == cudaSuccess;
    }();
    if (!attr_set) return ctx->fail(NZCB_E_CUDA, "witness: cannot reserve %zu bytes of shared memory", smem);
    static const int force = [] {
        const char* e = getenv("NZCB_WITNESS_CTA");  // 384 / 128 / 1284 (128 threads, 4 per SM) / 64: pin one variant (measurements)
        return e ? atoi(e) : 0;
    }();
    const bool batch = force ? force != (int)WIT_THREADS : B > (size_t)ctx->sm_count;
    if (batch && force == 64) {
        const uint32_t grid = (uint32_t)std::min<size_t>(B, (size_t)ctx->sm_count * 6);
        NZ_LAUNCH(ctx, (k_witness<64, 6>), grid, 64, smem_64, pv, d_inputs, d_wires, d_status, (uint32_t)B);
    } else if (batch && force == 1284) {
        const uint32_t grid = (uint32_t)std::min<size_t>(B, (size_t)ctx->sm_count * 4);
        NZ_LAUNCH(ctx, (k_witness<WIT_THREADS_BATCH, 4>), grid, WIT_THREADS_BATCH, smem_b, pv, d_inputs, d_wires, d_status, (uint32_t)B);
    } else if (batch) {
        const uint32_t grid = (uint32_t)std::min<size_t>(B, (size_t)ctx->sm_count * 3);
        NZ_LAUNCH(ctx, (k_witness<WIT_THREADS_BATCH, 3>), grid, WIT_THREADS_BATCH, smem_b, pv, d_inputs, d_wires, d_status, (uint32_t)B);
    } else {
        const uint32_t grid = (uint32_t)std::min<size_t>(B, (size_t)ctx->sm_count * 8);
        NZ_LAUNCH(ctx, (k_witness<WIT_THREADS, 1>), grid, WIT_THREADS, smem, pv, d_inputs, d_wires, d_status, (uint32_t)B);
    }
    return 0;
}
uint32_t circuit_n_total(const nzcb_circuit* c) { return c->n_total; }
uint32_t circuit_n_witness(const nzcb_circuit* c) { return c->n_witness; }
uint32_t circuit_n_in(const nzcb_circuit* c) { return c->n_in; }
}  // namespace nzcb

// passes per launch of a large batch: as many as ~24 GiB of wires hold, rounded down to whole waves of the batch kernel
static size_t witness_chunk(const nzcb_ctx* ctx, size_t per_pass, size_t B) {
    size_t chunk = std::max<size_t>(1, ((size_t)24 << 30) / per_pass);
    const size_t wave = (size_t)ctx->sm_count * 3;
    if (chunk > wave) chunk -= chunk % wave;
    return std::min(chunk, B);
}

extern "C" int32_t nzcb_witness_batch(nzcb_ctx* ctx, const nzcb_circuit* c, const uint8_t* inputs_le, size_t B,
                                      uint8_t* wtns_out, int32_t* status) {
    if (!ctx || !c || (!inputs_le && c->n_in) || !status) return NZCB_E_INVALID;
    if (c->ctx != ctx) return ctx->fail(NZCB_E_INVALID, "circuit was loaded on a different context");
    if (B == 0) return 0;
    NZ_CUDA(ctx, cudaSetDevice(ctx->device));
    // bound the device footprint: process the batch in chunks of at most ~24 GiB of wires, a whole number of waves
    // of the batch kernel (3 CTAs per SM) so that no launch leaves CTA slots empty
    const size_t per_pass = (size_t)c->n_total * sizeof(Fr);
    const size_t chunk = witness_chunk(ctx, per_pass, B);
    Fr* d_w = (Fr*)ctx->scratch_get("wt_wires", chunk * per_pass);
    Fr* d_in = (Fr*)ctx->scratch_get("wt_inputs", std::max<size_t>(32, chunk * (size_t)c->n_in * sizeof(Fr)));
    int32_t* d_st = (int32_t*)ctx->scratch_get("wt_status", chunk * sizeof(int32_t));
    if (!d_w || !d_in || !d_st) return ctx->fail(NZCB_E_NOMEM, "witness: cannot allocate %zu device bytes", chunk * per_pass);
    NZ_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
    for (size_t done = 0; done < B; done += chunk) {
        const size_t nb = std::min(chunk, B - done);
        if (c->n_in)
            NZ_CUDA(ctx, cudaMemcpyAsync(d_in, inputs_le + done * (size_t)c->n_in * 32, nb * (size_t)c->n_in * 32,
                                         cudaMemcpyHostToDevice, ctx->stream));
        NZ_TRY(witness_dev(ctx, c, d_in, nb, d_w, d_st));
        NZ_CUDA(ctx, cudaMemcpyAsync(status + done, d_st, nb * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
        if (wtns_out) {
            if (c->n_total == c->n_witness) {
                NZ_CUDA(ctx, cudaMemcpyAsync(wtns_out + done * (size_t)c->n_witness * 32, d_w, nb * per_pass,
                                             cudaMemcpyDeviceToHost, ctx->stream));
            } else {
                NZ_CUDA(ctx, cudaMemcpy2DAsync(wtns_out + done * (size_t)c->n_witness * 32, (size_t)c->n_witness * 32, d_w,
                                               per_pass, (size_t)c->n_witness * 32, nb, cudaMemcpyDeviceToHost,
                                               ctx->stream));
            }
        }
        NZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    NZ_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
    NZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    cudaEventElapsedTime(&ctx->last_ms, ctx->ev0, ctx->ev1);
    return 0;
}

// ---- large batches: what 65,536 passes can afford to bring back (BASELINE.json configs[3]) ----------------------
namespace {
__device__ __forceinline__ uint64_t splitmix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
// digest[pass] = sum_{i < n_witness, k < 8} (w_i[k] + 1) * splitmix64(8 i + k)  mod 2^64 -- position dependent, order free
__global__ void __launch_bounds__(256) k_witness_digest(const Fr* __restrict__ wires, uint32_t n_total, uint32_t n_witness,
                                                        uint64_t* __restrict__ digest) {
    __shared__ uint64_t part[8];
    const Fr* W = wires + (size_t)blockIdx.x * n_total;
    uint64_t acc = 0;
    for (uint32_t i = threadIdx.x; i < n_witness; i += blockDim.x) {
        const Fr w = W[i];
#pragma unroll
        for (uint32_t k = 0; k < 8; k++) acc += ((uint64_t)w.v[k] + 1) * splitmix64(8ull * i + k);
    }
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, off);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint64_t t = 0;
        for (int w = 0; w < 8; w++) t += part[w];
        digest[blockIdx.x] = t;
    }
}
}  // namespace

static int32_t witness_batch_ex_impl(nzcb_ctx* ctx, const nzcb_circuit* c, const uint8_t* inputs_le, bool inputs_on_device,
                                     size_t B, uint8_t* outputs_le, uint64_t* digest, size_t sample_stride,
                                     uint8_t* wtns_sample_out, int32_t* status);

extern "C" int32_t nzcb_witness_batch_ex(nzcb_ctx* ctx, const nzcb_circuit* c, const uint8_t* inputs_le, size_t B,
                                         uint8_t* outputs_le, uint64_t* digest, size_t sample_stride,
                                         uint8_t* wtns_sample_out, int32_t* status) {
    return witness_batch_ex_impl(ctx, c, inputs_le, false, B, outputs_le, digest, sample_stride, wtns_sample_out, status);
}
extern "C" int32_t nzcb_witness_batch_ex_dev(nzcb_ctx* ctx, const nzcb_circuit* c, const void* d_inputs_le, size_t B,
                                             uint8_t* outputs_le, uint64_t* digest, size_t sample_stride,
                                             uint8_t* wtns_sample_out, int32_t* status) {
    return witness_batch_ex_impl(ctx, c, (const uint8_t*)d_inputs_le, true, B, outputs_le, digest, sample_stride,
                                 wtns_sample_out, status);
}

static int32_t witness_batch_ex_impl(nzcb_ctx* ctx, const nzcb_circuit* c, const uint8_t* inputs_le, bool inputs_on_device,
                                     size_t B, uint8_t* outputs_le, uint64_t* digest, size_t sample_stride,
                                     uint8_t* wtns_sample_out, int32_t* status) {
    if (!ctx || !c || (!inputs_le && c->n_in) || !status || (wtns_sample_out && sample_stride == 0)) return NZCB_E_INVALID;
    if (c->ctx != ctx) return ctx->fail(NZCB_E_INVALID, "circuit was loaded on a different context");
    if (B == 0) return 0;
    NZ_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t per_pass = (size_t)c->n_total * sizeof(Fr);
    const size_t chunk = witness_chunk(ctx, per_pass, B);
    Fr* d_w = (Fr*)ctx->scratch_get("wt_wires", chunk * per_pass);
    const size_t in_bytes = std::max<size_t>(32, chunk * (size_t)c->n_in * sizeof(Fr));
    Fr* d_in2[2] = {(Fr*)ctx->scratch_get("wt_inputs", in_bytes), nullptr};
    int32_t* d_st = (int32_t*)ctx->scratch_get("wt_status", chunk * sizeof(int32_t));
    uint64_t* d_dg = (uint64_t*)ctx->scratch_get("wt_digest", chunk * sizeof(uint64_t));
    if (!d_w || !d_in2[0] || !d_st || !d_dg) return ctx->fail(NZCB_E_NOMEM, "witness: cannot allocate %zu device bytes", chunk * per_pass);
    // Host inputs of a batch longer than one chunk: chunk k + 1 is staged (host memcpy into pinned memory, DMA on a copy
    // stream into the second input buffer) while chunk k's kernel runs; a single chunk is copied straight from the caller.
    const bool staged = !inputs_on_device && c->n_in && B > chunk;
    if (staged) {
        d_in2[1] = (Fr*)ctx->scratch_get("wt_inputs_b", in_bytes);
        if (!d_in2[1]) return ctx->fail(NZCB_E_NOMEM, "witness: cannot allocate the second input buffer");
        if (ctx->wt_pin_bytes < in_bytes) {
            for (int b = 0; b < 2; b++) {
                if (ctx->wt_pin[b]) cudaFreeHost(ctx->wt_pin[b]);
                ctx->wt_pin[b] = nullptr;
            }
            ctx->wt_pin_bytes = 0;
            if (cudaHostAlloc(&ctx->wt_pin[0], in_bytes, cudaHostAllocDefault) != cudaSuccess ||
                cudaHostAlloc(&ctx->wt_pin[1], in_bytes, cudaHostAllocDefault) != cudaSuccess) {
                cudaGetLastError();
                return ctx->fail(NZCB_E_NOMEM, "witness: cannot allocate %zu bytes of pinned staging memory", 2 * in_bytes);
            }
            ctx->wt_pin_bytes = in_bytes;
        }
        if (!ctx->wt_copy) NZ_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->wt_copy, cudaStreamNonBlocking));
        for (int b = 0; b < 2; b++)
            if (!ctx->wt_ev[b]) NZ_CUDA(ctx, cudaEventCreateWithFlags(&ctx->wt_ev[b], cudaEventDisableTiming));
    }
    // stage(k): inputs of chunk k -> pinned[k & 1] -> d_in2[k & 1] on the copy stream; the event marks its arrival.
    // Both buffers of parity k & 1 were last used by chunk k - 2, whose kernel and copy are complete (see the loop).
    auto stage = [&](size_t k) -> int {
        const size_t lo = k * chunk, nbk = std::min(chunk, B - lo), bytes = nbk * (size_t)c->n_in * 32;
        const int b = (int)(k & 1);
        memcpy(ctx->wt_pin[b], inputs_le + lo * (size_t)c->n_in * 32, bytes);
        NZ_CUDA(ctx, cudaMemcpyAsync(d_in2[b], ctx->wt_pin[b], bytes, cudaMemcpyHostToDevice, ctx->wt_copy));
        NZ_CUDA(ctx, cudaEventRecord(ctx->wt_ev[b], ctx->wt_copy));
        return 0;
    };
    NZ_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
    if (staged) NZ_TRY(stage(0));
    for (size_t done = 0, k = 0; done < B; done += chunk, k++) {
        const size_t nb = std::min(chunk, B - done);
        const Fr* cur_in = d_in2[0];
        if (inputs_on_device) cur_in = reinterpret_cast<const Fr*>(inputs_le) + done * (size_t)c->n_in;
        else if (staged) {
            cur_in = d_in2[k & 1];
            NZ_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->wt_ev[k & 1], 0));
        } else if (c->n_in)
            NZ_CUDA(ctx, cudaMemcpyAsync(d_in2[0], inputs_le + done * (size_t)c->n_in * 32, nb * (size_t)c->n_in * 32,
                                         cudaMemcpyHostToDevice, ctx->stream));
        NZ_TRY(witness_dev(ctx, c, cur_in, nb, d_w, d_st));
        if (digest) NZ_LAUNCH(ctx, k_witness_digest, (unsigned)nb, 256, 0, d_w, c->n_total, c->n_witness, d_dg);
        // the next chunk's inputs travel while this chunk's kernel runs: staged BEFORE the read-backs below, which block
        // the host until the kernel is done (device-to-pageable copies are synchronous for the host).  Its buffers were
        // last used by chunk k - 1, complete since the previous iteration's synchronisation.
        if (staged && done + chunk < B) NZ_TRY(stage(k + 1));
        NZ_CUDA(ctx, cudaMemcpyAsync(status + done, d_st, nb * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
        if (digest)
            NZ_CUDA(ctx, cudaMemcpyAsync(digest + done, d_dg, nb * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
        if (outputs_le && c->n_out)
            NZ_CUDA(ctx, cudaMemcpy2DAsync(outputs_le + done * (size_t)c->n_out * 32, (size_t)c->n_out * 32, d_w + 1, per_pass,
                                           (size_t)c->n_out * 32, nb, cudaMemcpyDeviceToHost, ctx->stream));
        if (wtns_sample_out)
            for (size_t i = (done + sample_stride - 1) / sample_stride * sample_stride; i < done + nb; i += sample_stride)
                NZ_CUDA(ctx, cudaMemcpyAsync(wtns_sample_out + (i / sample_stride) * (size_t)c->n_witness * 32,
                                             d_w + (i - done) * (size_t)c->n_total, (size_t)c->n_witness * 32,
                                             cudaMemcpyDeviceToHost, ctx->stream));
        NZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    NZ_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
    NZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    cudaEventElapsedTime(&ctx->last_ms, ctx->ev0, ctx->ev1);
    return 0;
}
