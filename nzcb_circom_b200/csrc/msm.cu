// BN254 G1 multi-scalar multiplication -- replaces ffjavascript
// G1.multiExpAffine (un-vendored, /root/reference/yarn.lock:3905; called nine
// times per proof by snarkjs plonk.prove).
//
// Pippenger with signed c-bit digits (buckets 1..2^(c-1); a negative digit adds
// the negated base; scalars above (r-1)/2 are replaced by r - s with the sign
// flipped, so small negative wire values cost as little as small positive ones).
//
// Two bucket layouts:
//   table mode   (fixed bases, the zkey's SRS and its Lagrange form): the bases come with
//                the window shifts 2^(c w) P_i precomputed once (G1Table), so every window
//                feeds ONE bucket set per MSM and there is no per-window Horner;
//                c = 20 -> 13 digits per scalar instead of 16.
//   window mode  (one-shot bases, nzcb_msm_g1): one bucket set per window, Horner
//                over the windows at the end.
// Several MSMs over the same bases (A/B/C, T1/T2/T3, Wxi/Wxiw) run as one batch of
// "jobs": one sort, one accumulation launch, one reduction.
//
//   1. k_msm_digits<COUNT>   recode, histogram the bucket keys (warp-aggregated atomics)
//   2. scan_excl             bucket offsets (multi-block)
//   3. k_msm_digits<SCATTER> counting-sort the point references by bucket
//   4. k_msm_accum           THE hot kernel.  Chunks of 64 consecutive entries of the sorted
//                            list per thread, tiles of 256 chunks drawn from an atomic counter by
//                            a persistent grid: XYZZ += affine (8M + 2S) along the chunk;
//                            buckets inside a chunk are written directly, the first / last
//                            (possibly shared with the neighbours) go to a (key, partial) list
//      k_seg_join            adds up the short runs of that list (a bucket straddling a boundary)
//      k_seg_level           segmented reduction of what is left, level by level
//   5. k_bred                sum_b (b+1) B_b: chunked running sums over all buckets,
//      k_bred_pair           then one halving per launch
//      k_msm_horner          window mode: Horner over the windows
// Latency mode: a ctx may own only a slice of the point range (msm_table_finish exchanges the
// partial sums).   Algorithmic work: n * 16 * 10 modmul in step 4.
#include "common.cuh"
#include "msm_affine.cuh"
#include <dlfcn.h>
#include <stdlib.h>
#include <algorithm>
#include <vector>

namespace nzcb {

constexpr uint32_t KEY_NONE = 0xffffffffu;

struct MsmPlan {
    uint32_t c;      // window bits
    uint32_t W;      // digits per scalar
    uint32_t nbw;    // buckets per set = 2^(c-1)
    uint32_t G;      // bucket sets: K (table mode) or K * W (window mode)
    bool unified;    // table mode
    uint32_t stride; // table row stride (points)
    bool sparse;     // caller's hint: most scalars are tiny (round 1 in the Lagrange basis) -- XYZZ walk only
};

static uint32_t floor_log2(size_t n) {
    uint32_t lg = 0;
    while (((size_t)2 << lg) <= n) lg++;
    return lg;
}

static uint32_t env_window(uint32_t c, const char* name, uint32_t hi) {
    const char* env = getenv(name);
    if (env) {
        int v = atoi(env);
        if (v >= 2 && v <= (int)hi) c = (uint32_t)v;
    }
    return c;
}

uint32_t msm_table_window(size_t n) {
    int c = (int)floor_log2(n ? n : 1) - 1;
    c = std::max(4, std::min(20, c));
    return env_window((uint32_t)c, "NZCB_MSM_TABLE_WINDOW", 22);
}

static MsmPlan make_plan_window(size_t n, int K) {
    int c = (int)floor_log2(n ? n : 1) - 5;
    c = std::max(4, std::min(16, c));
    MsmPlan p;
    p.c = env_window((uint32_t)c, "NZCB_MSM_WINDOW", 20);
    p.W = 254 / p.c + 1;
    p.nbw = 1u << (p.c - 1);
    p.G = (uint32_t)K * p.W;
    p.unified = false;
    p.stride = 0;
    p.sparse = false;
    return p;
}

__device__ __forceinline__ uint32_t get_bits(const uint32_t* s, uint32_t off, uint32_t c) {
    const uint32_t limb = off >> 5, sh = off & 31;
    if (limb >= 8) return 0;
    uint32_t v = s[limb] >> sh;
    if (sh + c > 32 && limb + 1 < 8) v |= s[limb + 1] << (32 - sh);
    return v & ((1u << c) - 1);
}

constexpr int NZ_MSM_MAXJOBS = 4;
struct DigitArgs {
    const uint32_t* scalars[NZ_MSM_MAXJOBS];
    uint32_t lo[NZ_MSM_MAXJOBS];  // this ctx handles scalars [lo, n) of the job (latency mode: its slice)
    uint32_t n[NZ_MSM_MAXJOBS];
    uint32_t mont[NZ_MSM_MAXJOBS];
    uint32_t c, W, nbw, unified, stride;
};

// s > (r - 1) / 2 ?
__device__ __forceinline__ bool above_half(const Fr& s) {
    constexpr uint32_t H[8] = {0xf8000000u, 0xa1f0fac9u, 0x3cdcb848u, 0x9419f424u,
                               0x40c0ac2eu, 0xdc2822dbu, 0x7098d014u, 0x18322739u};
#pragma unroll
    for (int i = 7; i >= 0; i--) {
        if (s.v[i] > H[i]) return true;
        if (s.v[i] < H[i]) return false;
    }
    return false;
}

// COUNT: histogram into cnt.   !COUNT: scatter, cnt is the per-bucket cursor (zeroed), offsets the scan.
// Signed-digit recode of one scalar per thread; the digit loop is warp uniform so that lanes holding the SAME bucket
// key (wire values are mostly 0 / +-1 / bytes in round 1) combine their atomics: one atomicAdd per distinct key
// and warp, the lanes of a group take consecutive slots.
// AGG: warp-aggregated atomics (__match_any_sync) -- pays when many lanes hit the same bucket (wire values in the
// Lagrange basis); with uniform scalars and 2^19 buckets two lanes of a warp almost never meet, and the match
// instruction itself was most of the kernel's stall time (ncu: short_scoreboard 21, mio_throttle 12 per issue).
template <bool COUNT, bool AGG>
__global__ void __launch_bounds__(256) k_msm_digits(DigitArgs a, uint32_t* __restrict__ cnt,
                                                    const uint32_t* __restrict__ offsets, uint32_t* __restrict__ sorted) {
    const uint32_t k = blockIdx.y;
    const uint32_t i = a.lo[k] + blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31;
    bool live = i < a.n[k];
    Fr s = Fr::zero();
    if (live) {
        s = reinterpret_cast<const Fr*>(a.scalars[k])[i];
        if (a.mont[k]) {
            s = s.from_mont();
        } else {
            // multiExpAffine takes plain 256-bit integers: s * P = (s mod r) * P, and 2^256 < 6 r
            for (int it = 0; it < 5; it++) s = Fr::reduce_once(s);
        }
        live = !s.is_zero();
    }
    if (__ballot_sync(0xffffffffu, live) == 0) return;
    const bool neg = live && above_half(s);
    if (neg) s = Fr::modulus() - s;
    uint32_t carry = 0;
    const uint32_t half = 1u << (a.c - 1);
    if (!AGG) {
        // dense scalars: no lane cooperation, so four digits at a time -- their atomics (and the offset loads) are
        // in flight together before the first reference is stored: one memory round trip per four digits, not four
        for (uint32_t w0 = 0; w0 < a.W; w0 += 4) {
            uint32_t key[4], ref[4], base[4], off[4];
            bool has[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const uint32_t w = w0 + u;
                has[u] = false;
                if (w >= a.W) continue;
                uint32_t d = get_bits(s.v, w * a.c, a.c) + carry;
                bool dneg = false;
                if (d > half) {
                    carry = 1;
                    d = (1u << a.c) - d;
                    dneg = true;
                } else {
                    carry = 0;
                }
                has[u] = live && d != 0;
                const uint32_t set = a.unified ? k : k * a.W + w;
                key[u] = set * a.nbw + (d - 1);
                ref[u] = (a.unified ? w * a.stride + i : i) | ((neg != dneg) ? 0x80000000u : 0u);
            }
#pragma unroll
            for (int u = 0; u < 4; u++)
                if (has[u]) {
                    base[u] = atomicAdd(&cnt[key[u]], 1u);
                    if (!COUNT) off[u] = offsets[key[u]];
                }
            if (!COUNT) {
#pragma unroll
                for (int u = 0; u < 4; u++)
                    if (has[u]) sorted[off[u] + base[u]] = ref[u];
            }
        }
    }
    for (uint32_t w = 0; AGG && w < a.W; w++) {
        uint32_t d = get_bits(s.v, w * a.c, a.c) + carry;
        bool dneg = false;
        if (d > half) {
            carry = 1;
            d = (1u << a.c) - d;  // |d - 2^c|; 0 when d == 2^c (digit 0, carry 1)
            dneg = true;
        } else {
            carry = 0;
        }
        const bool has = live && d != 0;
        const uint32_t active = __ballot_sync(0xffffffffu, has);
        if (!has) continue;  // the lanes with a digit stay converged on `active`
        const uint32_t set = a.unified ? k : k * a.W + w;
        const uint32_t key = set * a.nbw + (d - 1);
        uint32_t base = 0, rank = 0;
        if (AGG) {
            const uint32_t peers = __match_any_sync(active, key);
            const uint32_t leader = __ffs(peers) - 1;
            rank = __popc(peers & ((1u << lane) - 1));
            if (lane == leader) base = atomicAdd(&cnt[key], (uint32_t)__popc(peers));
            if (!COUNT) base = __shfl_sync(peers, base, leader);
        } else {
            base = atomicAdd(&cnt[key], 1u);
        }
        if (!COUNT) {
            const uint32_t ref = a.unified ? w * a.stride + i : i;
            sorted[offsets[key] + base + rank] = ref | ((neg != dneg) ? 0x80000000u : 0u);
        }
    }
}

// ---- exclusive scan of n u32 counts into offsets[0..n], offsets[n] = total ------------------
constexpr uint32_t SCAN_ITEMS = 8, SCAN_THREADS = 256, SCAN_TILE = SCAN_ITEMS * SCAN_THREADS;

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_tile(const uint32_t* __restrict__ in, uint32_t* __restrict__ out,
                                                            uint32_t* __restrict__ tile_sums, size_t n) {
    __shared__ uint32_t part[SCAN_THREADS];
    const uint32_t t = threadIdx.x;
    const size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)t * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS], sum = 0;
#pragma unroll
    for (uint32_t j = 0; j < SCAN_ITEMS; j++) {
        v[j] = base + j < n ? in[base + j] : 0u;
        sum += v[j];
    }
    part[t] = sum;
    __syncthreads();
    for (uint32_t off = 1; off < SCAN_THREADS; off <<= 1) {
        const uint32_t x = t >= off ? part[t - off] : 0u;
        __syncthreads();
        part[t] += x;
        __syncthreads();
    }
    uint32_t run = part[t] - sum;
#pragma unroll
    for (uint32_t j = 0; j < SCAN_ITEMS; j++) {
        if (base + j < n) out[base + j] = run;
        run += v[j];
    }
    if (t == SCAN_THREADS - 1) tile_sums[blockIdx.x] = part[t];
}
// single block: exclusive scan of the tile sums in place; writes the grand total to *total
__global__ void __launch_bounds__(1024) k_scan_tops(uint32_t* __restrict__ sums, size_t n, uint32_t* __restrict__ total) {
    __shared__ uint32_t part[1024];
    const uint32_t t = threadIdx.x;
    const size_t per = (n + 1023) / 1024;
    const size_t lo = (size_t)t * per, hi = lo + per < n ? lo + per : n;
    uint32_t sum = 0;
    for (size_t k = lo; k < hi; k++) sum += sums[k];
    part[t] = sum;
    __syncthreads();
    for (uint32_t off = 1; off < 1024; off <<= 1) {
        const uint32_t x = t >= off ? part[t - off] : 0u;
        __syncthreads();
        part[t] += x;
        __syncthreads();
    }
    uint32_t run = part[t] - sum;
    for (size_t k = lo; k < hi; k++) {
        const uint32_t c = sums[k];
        sums[k] = run;
        run += c;
    }
    if (t == 1023) *total = part[1023];
}
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_add(uint32_t* __restrict__ out, const uint32_t* __restrict__ tile_offs,
                                                           size_t n) {
    const size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
    const uint32_t add = tile_offs[blockIdx.x];
#pragma unroll
    for (uint32_t j = 0; j < SCAN_ITEMS; j++)
        if (base + j < n) out[base + j] += add;
}

static int scan_excl(nzcb_ctx* ctx, const uint32_t* counts, uint32_t* offsets, size_t n) {
    const size_t tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    uint32_t* sums = (uint32_t*)ctx->scratch_get("msm_scan_sums", (tiles + 1) * 4);
    if (!sums) return ctx->fail(NZCB_E_NOMEM, "msm: cannot allocate scan workspace");
    NZ_LAUNCH(ctx, k_scan_tile, (unsigned)tiles, SCAN_THREADS, 0, counts, offsets, sums, n);
    NZ_LAUNCH(ctx, k_scan_tops, 1, 1024, 0, sums, tiles, offsets + n);
    NZ_LAUNCH(ctx, k_scan_add, (unsigned)tiles, SCAN_THREADS, 0, offsets, sums, n);
    return 0;
}

// ---- step 4: accumulation over equal-length chunks -------------------------------------------
// The sorted list is cut into chunks of L consecutive entries (L chosen on the host from the upper bound of the
// entry count, so the (key, partial) list has a host-known size); a tile = ACC_THREADS chunks.  The grid is
// persistent -- one CTA per resident slot -- and CTAs draw tiles from an atomic counter, so the kernel keeps its
// balance whatever shares the GPU with it (other lanes' kernels, the witness program).
// Per chunk: two (key, partial) slots for its first and last bucket (possibly shared with the neighbours);
// every other bucket it meets lies wholly inside the chunk and is written straight to buckets[].
constexpr uint32_t ACC_THREADS = 256, ACC_LMIN = 8, ACC_LMAX = 64;

// bucket holding the first entry of every chunk (binary search in the offsets, off the hot kernel's critical path).
// `shift`: the list has been through that many halving rounds (msm_affine.cuh), every offset is a multiple of 2^shift.
__global__ void __launch_bounds__(256) k_msm_chunk_buckets(const uint32_t* __restrict__ offsets, uint32_t n_keys, uint32_t L,
                                                           uint32_t shift, uint32_t* __restrict__ chunk_bucket,
                                                           uint32_t max_chunks) {
    const uint32_t ch = blockIdx.x * blockDim.x + threadIdx.x;
    if (ch >= max_chunks) return;
    const uint32_t E = offsets[n_keys] >> shift;
    const uint64_t lo = (uint64_t)ch * L;
    if (lo >= E) return;
    // the largest b with offsets[b] <= lo  (offsets[n_keys] = E > lo)
    uint32_t b = 0, z = n_keys;
    while (z - b > 1) {
        const uint32_t m = (b + z) >> 1;
        if ((offsets[m] >> shift) <= (uint32_t)lo) b = m;
        else z = m;
    }
    chunk_bucket[ch] = b;
}

// DIRECT: the list holds the points themselves (the output of the halving rounds), not references into `bases`.
template <bool DIRECT>
__global__ void __launch_bounds__(ACC_THREADS, 2)
    k_msm_accum(const G1Affine* __restrict__ bases, const uint32_t* __restrict__ sorted, const uint32_t* __restrict__ offsets,
                uint32_t n_keys, uint32_t L, uint32_t shift, const uint32_t* __restrict__ chunk_bucket,
                uint32_t* __restrict__ tile_counter, G1XYZZ* __restrict__ buckets, uint32_t* __restrict__ pkeys,
                G1XYZZ* __restrict__ pvals) {
    __shared__ uint32_t s_tile;
    const uint32_t E = offsets[n_keys] >> shift;
    const uint32_t n_chunks = (uint32_t)(((uint64_t)E + L - 1) / L);
    const uint32_t n_tiles = (n_chunks + ACC_THREADS - 1) / ACC_THREADS;
    for (;;) {
        if (threadIdx.x == 0) s_tile = atomicAdd(tile_counter, 1u);
        __syncthreads();
        const uint32_t tile = s_tile;
        __syncthreads();
        if (tile >= n_tiles) break;
        const uint32_t t = tile * ACC_THREADS + threadIdx.x;  // chunk id
        if (t >= n_chunks) continue;
        const uint32_t lo = t * L;  // < E <= 2^32 - 1
        const uint32_t hi = (uint64_t)lo + L < E ? lo + L : E;
        uint32_t b = chunk_bucket[t];
        uint32_t next = offsets[b + 1] >> shift;
        G1XYZZ acc = G1XYZZ::inf();
        bool first = true;
        uint32_t e = DIRECT ? 0u : sorted[lo];
        G1Affine p = DIRECT ? bases[lo] : bases[e & 0x7fffffffu];
        for (uint32_t k = lo; k < hi; k++) {
            // prefetch the next point while this one is added
            const uint32_t e_cur = e;
            const G1Affine p_cur = p;
            if (k + 1 < hi) {
                if (DIRECT) {
                    p = bases[k + 1];
                } else {
                    e = sorted[k + 1];
                    p = bases[e & 0x7fffffffu];
                }
            }
            if (k == next) {  // bucket boundary: flush
                if (first) {
                    pkeys[2 * t] = b;
                    pvals[2 * t] = acc;
                    first = false;
                } else {
                    buckets[b] = acc;
                }
                acc = G1XYZZ::inf();
                do {
                    b++;
                    next = offsets[b + 1] >> shift;
                } while (next == k);
            }
            G1Affine q = p_cur;
            if (!DIRECT && (e_cur & 0x80000000u)) q.y = q.y.neg();  // (0,0) stays (0,0)
            acc.add_affine(q);
        }
        if (first) {
            pkeys[2 * t] = b;
            pvals[2 * t] = acc;
            pkeys[2 * t + 1] = b;
            pvals[2 * t + 1] = G1XYZZ::inf();
        } else {
            pkeys[2 * t + 1] = b;
            pvals[2 * t + 1] = acc;
        }
    }
}

// ---- step 4a: halving rounds with batched affine additions (msm_affine.cuh) ---------------------
// counts -> multiples of 2^R, so that every bucket's segment of the sorted list is aligned for R rounds
__global__ void __launch_bounds__(256) k_pad_counts(uint32_t* __restrict__ counts, size_t n, uint32_t R) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t m = (1u << R) - 1;
    counts[i] = (counts[i] + m) & ~m;
}
// round `shift` (1-based): the input list has (*e_total >> (shift - 1)) entries, the output half as many
template <bool REFS>
__global__ void __launch_bounds__(256) k_aff_forward(const G1Affine* __restrict__ pts, const uint32_t* __restrict__ refs,
                                                     const uint32_t* __restrict__ e_total, uint32_t shift,
                                                     Fq* __restrict__ P, Fq* __restrict__ totals) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t n_add = *e_total >> shift;
    if (REFS) {
        const AffRefSrc src{pts, refs};
        aff_forward_body(t, src, n_add, P, totals);
    } else {
        const AffPtSrc src{pts};
        aff_forward_body(t, src, n_add, P, totals);
    }
}
__global__ void __launch_bounds__(128) k_aff_invert(Fq* __restrict__ totals, Fq* __restrict__ tmp,
                                                    const uint32_t* __restrict__ e_total, uint32_t shift) {
    const size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t n_add = *e_total >> shift;
    const size_t n_tot = (n_add + AFF_M - 1) / AFF_M;
    const size_t lo = s * AFF_INV_CHUNK;
    if (lo >= n_tot) return;
    aff_invert_chunk(totals, tmp, lo, lo + AFF_INV_CHUNK < n_tot ? lo + AFF_INV_CHUNK : n_tot);
}
template <bool REFS, int CTAS>
__global__ void __launch_bounds__(256, CTAS) k_aff_backward(const G1Affine* __restrict__ pts, const uint32_t* __restrict__ refs,
                                                         const uint32_t* __restrict__ e_total, uint32_t shift,
                                                         const Fq* __restrict__ P, const Fq* __restrict__ totals_inv,
                                                         G1Affine* __restrict__ out) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t n_add = *e_total >> shift;
    if (REFS) {
        const AffRefSrc src{pts, refs};
        aff_backward_body(t, src, n_add, P, totals_inv, out);
    } else {
        const AffPtSrc src{pts};
        aff_backward_body(t, src, n_add, P, totals_inv, out);
    }
}

// One level of the segmented reduction of a key-sorted (key, partial) list: thread j sums runs of equal keys in
// [j*L, (j+1)*L); runs strictly inside are complete -> buckets[key]; its first and last run go to the next list.
// final != 0: single thread, everything is written to buckets.
__global__ void __launch_bounds__(128) k_seg_level(const uint32_t* __restrict__ keys, const G1XYZZ* __restrict__ vals,
                                                   uint32_t N, uint32_t L, G1XYZZ* __restrict__ buckets,
                                                   uint32_t* __restrict__ okeys, G1XYZZ* __restrict__ ovals, int final) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t lo64 = (uint64_t)j * L;
    if (lo64 >= N) return;
    const uint32_t lo = (uint32_t)lo64;
    const uint32_t hi = (uint64_t)lo + L < N ? lo + L : N;
    uint32_t key = keys[lo];
    G1XYZZ acc = key != KEY_NONE ? vals[lo] : G1XYZZ::inf();
    bool first = !final;
    for (uint32_t k = lo + 1; k < hi; k++) {
        const uint32_t kk = keys[k];
        if (kk != key) {
            if (first) {
                okeys[2 * j] = key;
                ovals[2 * j] = acc;
                first = false;
            } else if (key != KEY_NONE) {
                buckets[key] = acc;
            }
            key = kk;
            acc = key != KEY_NONE ? vals[k] : G1XYZZ::inf();
        } else if (key != KEY_NONE) {
            acc.add(vals[k]);
        }
    }
    if (final) {
        if (key != KEY_NONE) buckets[key] = acc;
    } else if (first) {
        okeys[2 * j] = key;
        ovals[2 * j] = acc;
        okeys[2 * j + 1] = key;
        ovals[2 * j + 1] = G1XYZZ::inf();
    } else {
        okeys[2 * j + 1] = key;
        ovals[2 * j + 1] = acc;
    }
}

// Short runs first: almost every bucket that straddles a chunk boundary has two or three partials.  One thread per
// list entry; the thread at the head of a run of at most SEG_SHORT equal keys adds it up and completes the bucket.
// The keys of resolved entries are blanked in a second array, which the level-by-level reduction above then
// processes -- it only finds work when a bucket holds a large share of all points.
constexpr uint32_t SEG_SHORT = 6;
__global__ void __launch_bounds__(128) k_seg_join(const uint32_t* __restrict__ keys, const G1XYZZ* __restrict__ vals, uint32_t N,
                                                  G1XYZZ* __restrict__ buckets, uint32_t* __restrict__ okeys) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const uint32_t key = keys[i];
    if (key == KEY_NONE) {
        okeys[i] = KEY_NONE;
        return;
    }
    // position inside the run and run length, both capped just above SEG_SHORT
    uint32_t before = 0, after = 0;
    while (before <= SEG_SHORT && i > before && keys[i - before - 1] == key) before++;
    while (after <= SEG_SHORT && i + after + 1 < N && keys[i + after + 1] == key) after++;
    const bool is_short = before + after + 1 <= SEG_SHORT;
    okeys[i] = is_short ? KEY_NONE : key;
    if (!is_short || before != 0) return;
    G1XYZZ acc = vals[i];
    for (uint32_t k = 1; k <= after; k++) acc.add(vals[i + k]);
    buckets[key] = acc;
}

// ---- step 5: R(X) = sum_b weight(b) X_b per bucket set, weight = b + 1 (one_based) or b -------
// chunk ch of S = 2^log_S elements: run = sum X, acc = sum local_weight * X (+ the chunk's plain sums P);
// R(X) = sum_ch acc_ch + R0(S * run)  ->  Xo[ch] = S * run_ch (zero-based next level), Po[ch] = acc_ch + sum P.
__global__ void __launch_bounds__(128) k_bred(const G1XYZZ* __restrict__ X, const G1XYZZ* __restrict__ P, uint32_t total,
                                              uint32_t log_S, int one_based, G1XYZZ* __restrict__ Xo,
                                              G1XYZZ* __restrict__ Po) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const uint32_t S = 1u << log_S;
    const G1XYZZ* x = X + (size_t)t * S;
    G1XYZZ run = G1XYZZ::inf(), acc = G1XYZZ::inf();
    for (int k = (int)S - 1; k >= 1; k--) {
        run.add(x[k]);
        acc.add(run);
    }
    run.add(x[0]);
    if (one_based) acc.add(run);
    if (P) {
        const G1XYZZ* pp = P + (size_t)t * S;
        G1XYZZ pl = pp[0];
        for (uint32_t k = 1; k < S; k++) pl.add(pp[k]);
        acc.add(pl);
    }
    Po[t] = acc;
    for (uint32_t i = 0; i < log_S; i++) run = run.dbl();
    Xo[t] = run;
}

// The later levels of the same recursion are small and latency bound (one point addition is ~10 us of dependent
// multiplies), so they halve the array per launch with the two independent chains of a pair on different warps:
// the first half of the grid computes Xo[i] = 2 (X[2i] + X[2i+1]), the second Po[i] = P[2i] + P[2i+1] + X[2i+1]
// (zero-based weights: 0 * X[2i] + 1 * X[2i+1]).
__global__ void __launch_bounds__(128) k_bred_pair(const G1XYZZ* __restrict__ X, const G1XYZZ* __restrict__ P, uint32_t pairs,
                                                   uint32_t role_blocks, G1XYZZ* __restrict__ Xo, G1XYZZ* __restrict__ Po) {
    const bool second = blockIdx.x >= role_blocks;
    const uint32_t i = (second ? blockIdx.x - role_blocks : blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= pairs) return;
    if (!second) {
        G1XYZZ r = X[2 * (size_t)i];
        r.add(X[2 * (size_t)i + 1]);
        Xo[i] = r.dbl();
    } else {
        G1XYZZ q = P[2 * (size_t)i];
        q.add(P[2 * (size_t)i + 1]);
        q.add(X[2 * (size_t)i + 1]);
        Po[i] = q;
    }
}

// The last levels of the same recursion in ONE launch: one CTA per bucket set walks the remaining `bits` halvings
// with a barrier between levels instead of a launch (a level is one point addition deep: ~10 us of dependent
// multiplies against ~70 us per launch of a nearly empty grid).  The two roles of k_bred_pair share the CTA.
constexpr uint32_t BRED_TAIL_BITS = 9, BRED_TAIL_THREADS = 512;
__global__ void __launch_bounds__(BRED_TAIL_THREADS) k_bred_tail(G1XYZZ* Xa, G1XYZZ* Pa, G1XYZZ* Xb, G1XYZZ* Pb, uint32_t bits,
                                                                G1XYZZ* __restrict__ out) {
    const size_t base = (size_t)blockIdx.x << bits;
    Xa += base; Pa += base; Xb += base; Pb += base;
    for (uint32_t b = bits; b > 0; b--) {
        const uint32_t pairs = 1u << (b - 1);
        for (uint32_t i = threadIdx.x; i < 2 * pairs; i += blockDim.x) {
            if (i < pairs) {
                G1XYZZ r = Xa[2 * i];
                r.add(Xa[2 * i + 1]);
                Xb[i] = r.dbl();
            } else {
                const uint32_t j = i - pairs;
                G1XYZZ q = Pa[2 * j];
                q.add(Pa[2 * j + 1]);
                q.add(Xa[2 * j + 1]);
                Pb[j] = q;
            }
        }
        __syncthreads();
        G1XYZZ* t = Xa; Xa = Xb; Xb = t;
        t = Pa; Pa = Pb; Pb = t;
    }
    if (threadIdx.x == 0) out[blockIdx.x] = Pa[0];
}

// window mode: job k's result = sum_w 2^(c w) * set[k*W + w]   (Horner, one thread per job)
__global__ void k_msm_horner(const G1XYZZ* __restrict__ sets, uint32_t W, uint32_t c, G1XYZZ* __restrict__ out) {
    const uint32_t k = blockIdx.x;
    if (threadIdx.x != 0) return;
    G1XYZZ r = G1XYZZ::inf();
    for (int w = (int)W - 1; w >= 0; w--) {
        for (uint32_t i = 0; i < c; i++) r = r.dbl();
        r.add(sets[(size_t)k * W + w]);
    }
    out[k] = r;
}
__global__ void k_copy_pts(const G1XYZZ* __restrict__ in, G1XYZZ* __restrict__ out, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i];
}

__global__ void k_set_inf(G1XYZZ* out, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = G1XYZZ::inf();
}

struct MsmJob {
    const uint32_t* scalars;
    size_t lo, n;  // scalars [lo, n)
    bool mont;
};

static int msm_run(nzcb_ctx* ctx, const G1Affine* d_bases, const MsmPlan& p, const MsmJob* jobs, int K, G1XYZZ* d_out) {
    size_t n_max = 0, n_sum = 0;
    DigitArgs da;
    memset(&da, 0, sizeof(da));
    for (int k = 0; k < K; k++) {
        if (jobs[k].n >= ((size_t)1 << 26)) return ctx->fail(NZCB_E_INVALID, "msm: n too large");
        da.scalars[k] = jobs[k].scalars;
        da.lo[k] = (uint32_t)jobs[k].lo;
        da.n[k] = (uint32_t)jobs[k].n;
        da.mont[k] = jobs[k].mont ? 1 : 0;
        n_max = std::max(n_max, jobs[k].n - jobs[k].lo);
        n_sum += jobs[k].n - jobs[k].lo;
    }
    da.c = p.c; da.W = p.W; da.nbw = p.nbw; da.unified = p.unified ? 1 : 0; da.stride = p.stride;
    const size_t n_keys = (size_t)p.G * p.nbw;
    if (n_sum == 0) {
        NZ_LAUNCH(ctx, k_set_inf, 1, 32, 0, d_out, (uint32_t)K);
        return 0;
    }
    if (n_keys >= ((size_t)1 << 31)) return ctx->fail(NZCB_E_INVALID, "msm: batch too large");

    // halving rounds with batched affine additions first (msm_affine.cuh) when the buckets are full enough for the
    // padding of their segments to multiples of 2^R to be cheap; the XYZZ walk finishes (or does everything, R = 0)
    const size_t e_raw_max = n_sum * p.W;
    // (... and the list long enough: a round is three launches, one of them a 380-multiply Fermat chain of ~0.3 ms
    // whatever the size -- below ~2^24 entries the XYZZ walk alone is faster: primitive sweep, profiles/)
    uint32_t R = (!p.sparse && e_raw_max >= 24 * n_keys && e_raw_max >= ((size_t)1 << 24)) ? 3 : 0;
    {
        const char* env = getenv("NZCB_MSM_AFFINE");
        if (env && env[0] >= '0' && env[0] <= '5' && !env[1]) R = (uint32_t)(env[0] - '0');
    }
    const size_t e_max = e_raw_max + (((size_t)1 << R) - 1) * n_keys;  // upper bound of the padded list length
    if (e_max >= ((size_t)1 << 32)) return ctx->fail(NZCB_E_INVALID, "msm: batch too large");
    const size_t e_tail_max = (e_max >> R) + 1;                        // what the XYZZ walk sees

    // accumulation grid: persistent, one CTA per resident slot; chunk length from the entry-count upper bound
    static const int blocks_per_sm = [] {
        int v = 0, w = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, k_msm_accum<false>, ACC_THREADS, 0) != cudaSuccess || v < 1) v = 1;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&w, k_msm_accum<true>, ACC_THREADS, 0) != cudaSuccess || w < 1) w = 1;
        return std::min(v, w);
    }();
    const uint32_t acc_slots = (uint32_t)ctx->sm_count * (uint32_t)blocks_per_sm;
    uint32_t L = (uint32_t)std::min<size_t>(ACC_LMAX, std::max<size_t>(ACC_LMIN, e_tail_max / ((size_t)acc_slots * ACC_THREADS * 4)));
    {
        const char* env = getenv("NZCB_MSM_CHUNK");
        if (env && atoi(env) >= 1 && atoi(env) <= 4096) L = (uint32_t)atoi(env);
    }
    const uint32_t max_chunks = (uint32_t)((e_tail_max + L - 1) / L);
    const uint32_t acc_blocks = std::max<uint32_t>(1, std::min<uint32_t>(acc_slots, div_up(max_chunks, ACC_THREADS)));
    const uint32_t T1 = max_chunks;  // (key, partial) list: two slots per chunk

    uint32_t* counts = (uint32_t*)ctx->scratch_get("msm_counts", (n_keys + 1) * 4);
    uint32_t* offsets = (uint32_t*)ctx->scratch_get("msm_offsets", (n_keys + 1) * 4);
    uint32_t* sorted = (uint32_t*)ctx->scratch_get("msm_sorted", e_max * 4);
    G1XYZZ* buckets = (G1XYZZ*)ctx->scratch_get("msm_buckets", n_keys * sizeof(G1XYZZ));
    uint32_t* pk0 = (uint32_t*)ctx->scratch_get("msm_pk0", (size_t)2 * T1 * 4);
    uint32_t* pk1 = (uint32_t*)ctx->scratch_get("msm_pk1", (size_t)2 * T1 * 4);
    G1XYZZ* pv0 = (G1XYZZ*)ctx->scratch_get("msm_pv0", (size_t)2 * T1 * sizeof(G1XYZZ));
    G1XYZZ* pv1 = (G1XYZZ*)ctx->scratch_get("msm_pv1", (size_t)2 * T1 * sizeof(G1XYZZ));
    uint32_t* chunk_bucket = (uint32_t*)ctx->scratch_get("msm_chunk_bucket", (size_t)max_chunks * 4 + 4);
    uint32_t* tile_counter = (uint32_t*)ctx->scratch_get("msm_tile_counter", 256);
    // reduction ping-pong: level 1 output has n_keys / 2^log_S elements
    const size_t red_cap = std::max<size_t>(n_keys / 2, p.G) + 1;
    G1XYZZ* rx0 = (G1XYZZ*)ctx->scratch_get("msm_rx0", red_cap * sizeof(G1XYZZ));
    G1XYZZ* rp0 = (G1XYZZ*)ctx->scratch_get("msm_rp0", red_cap * sizeof(G1XYZZ));
    G1XYZZ* rx1 = (G1XYZZ*)ctx->scratch_get("msm_rx1", red_cap * sizeof(G1XYZZ));
    G1XYZZ* rp1 = (G1XYZZ*)ctx->scratch_get("msm_rp1", red_cap * sizeof(G1XYZZ));
    if (!counts || !offsets || !sorted || !buckets || !pk0 || !pk1 || !pv0 || !pv1 || !rx0 || !rp0 || !rx1 || !rp1 ||
        !chunk_bucket || !tile_counter)
        return ctx->fail(NZCB_E_NOMEM, "msm: cannot allocate workspace for %zu scalars", n_sum);
    // halving rounds: prefix products, batch totals, two point lists (round r reads one, writes the other)
    Fq *aff_P = nullptr, *aff_tot = nullptr, *aff_tmp = nullptr;
    G1Affine* aff_pts[2] = {nullptr, nullptr};
    if (R) {
        const size_t a1 = e_max / 2 + 1, t1 = a1 / AFF_M + 2;
        aff_P = (Fq*)ctx->scratch_get("msm_aff_P", a1 * sizeof(Fq));
        aff_tot = (Fq*)ctx->scratch_get("msm_aff_tot", t1 * sizeof(Fq));
        aff_tmp = (Fq*)ctx->scratch_get("msm_aff_tmp", t1 * sizeof(Fq));
        aff_pts[0] = (G1Affine*)ctx->scratch_get("msm_aff_pts0", a1 * sizeof(G1Affine));
        aff_pts[1] = (G1Affine*)ctx->scratch_get("msm_aff_pts1", (a1 / 2 + 1) * sizeof(G1Affine));
        if (!aff_P || !aff_tot || !aff_tmp || !aff_pts[0] || !aff_pts[1])
            return ctx->fail(NZCB_E_NOMEM, "msm: cannot allocate the affine-round workspace for %zu scalars", n_sum);
    }

    // 1-3: sort the point references by bucket
    NZ_CUDA(ctx, cudaMemsetAsync(counts, 0, (n_keys + 1) * 4, ctx->stream));
    const dim3 dgrid(div_up(n_max, 256), (unsigned)K);
    // aggregate the atomics of a warp when buckets are few or the scalars are known to be wire-like
    const bool agg = p.sparse || n_keys < ((size_t)1 << 16);
    if (agg) NZ_LAUNCH(ctx, (k_msm_digits<true, true>), dgrid, 256, 0, da, counts, nullptr, nullptr);
    else NZ_LAUNCH(ctx, (k_msm_digits<true, false>), dgrid, 256, 0, da, counts, nullptr, nullptr);
    if (R) {
        NZ_LAUNCH(ctx, k_pad_counts, div_up(n_keys, 256), 256, 0, counts, n_keys, R);
        NZ_CUDA(ctx, cudaMemsetAsync(sorted, 0xff, e_max * 4, ctx->stream));  // AFF_NULL in the padding slots
    }
    NZ_TRY(scan_excl(ctx, counts, offsets, n_keys));
    NZ_CUDA(ctx, cudaMemsetAsync(counts, 0, (n_keys + 1) * 4, ctx->stream));
    if (agg) NZ_LAUNCH(ctx, (k_msm_digits<false, true>), dgrid, 256, 0, da, counts, offsets, sorted);
    else NZ_LAUNCH(ctx, (k_msm_digits<false, false>), dgrid, 256, 0, da, counts, offsets, sorted);
    NZ_CUDA(ctx, cudaMemsetAsync(buckets, 0, n_keys * sizeof(G1XYZZ), ctx->stream));  // ZZ = 0: infinity
    NZ_CUDA(ctx, cudaMemsetAsync(pk0, 0xff, (size_t)2 * T1 * 4, ctx->stream));            // KEY_NONE beyond the last chunk
    NZ_CUDA(ctx, cudaMemsetAsync(tile_counter, 0, 4, ctx->stream));
    NZ_LAUNCH(ctx, k_msm_chunk_buckets, div_up(max_chunks, 256), 256, 0, offsets, (uint32_t)n_keys, L, R, chunk_bucket, max_chunks);

    // 4: accumulate
    if (ctx->prof_on) {
        if (ctx->prof_used == ctx->prof_ev.size()) {
            cudaEvent_t a, b;
            NZ_CUDA(ctx, cudaEventCreate(&a));
            NZ_CUDA(ctx, cudaEventCreate(&b));
            ctx->prof_ev.push_back({a, b});
        }
        if (ctx->prof_entries.size() < 4096) ctx->prof_entries.resize(4096, 0);  // fixed storage: the copies below land here
        if (ctx->prof_used < ctx->prof_entries.size())
            NZ_CUDA(ctx, cudaMemcpyAsync(&ctx->prof_entries[ctx->prof_used], offsets + n_keys, 4, cudaMemcpyDeviceToHost, ctx->stream));
        NZ_CUDA(ctx, cudaEventRecord(ctx->prof_ev[ctx->prof_used].first, ctx->stream));
    }
    const uint32_t* e_total = offsets + n_keys;
    static const int bwd_ctas = [] {   // resident CTAs per SM of the backward pass: 3 (80 registers) is 4 % faster alone
        const char* e = getenv("NZCB_AFF_BWD_CTAS");
        return e && atoi(e) == 2 ? 2 : 3;
    }();
    for (uint32_t r = 1; r <= R; r++) {
        const size_t adds = (e_max >> r) + 1;           // upper bound; the kernels read the count from e_total
        const unsigned fb = div_up(div_up(adds, AFF_M), 256);
        const unsigned ib = div_up(div_up(div_up(adds, AFF_M), AFF_INV_CHUNK), 128);
        G1Affine* out = aff_pts[(r - 1) & 1];
        if (r == 1) {
            NZ_LAUNCH(ctx, k_aff_forward<true>, fb, 256, 0, d_bases, sorted, e_total, r, aff_P, aff_tot);
            NZ_LAUNCH(ctx, k_aff_invert, ib, 128, 0, aff_tot, aff_tmp, e_total, r);
            if (bwd_ctas == 2) NZ_LAUNCH(ctx, (k_aff_backward<true, 2>), fb, 256, 0, d_bases, sorted, e_total, r, aff_P, aff_tot, out);
            else NZ_LAUNCH(ctx, (k_aff_backward<true, 3>), fb, 256, 0, d_bases, sorted, e_total, r, aff_P, aff_tot, out);
        } else {
            const G1Affine* in = aff_pts[r & 1];
            NZ_LAUNCH(ctx, k_aff_forward<false>, fb, 256, 0, in, nullptr, e_total, r, aff_P, aff_tot);
            NZ_LAUNCH(ctx, k_aff_invert, ib, 128, 0, aff_tot, aff_tmp, e_total, r);
            if (bwd_ctas == 2) NZ_LAUNCH(ctx, (k_aff_backward<false, 2>), fb, 256, 0, in, nullptr, e_total, r, aff_P, aff_tot, out);
            else NZ_LAUNCH(ctx, (k_aff_backward<false, 3>), fb, 256, 0, in, nullptr, e_total, r, aff_P, aff_tot, out);
        }
    }
    if (R) {
        NZ_LAUNCH(ctx, k_msm_accum<true>, acc_blocks, ACC_THREADS, 0, aff_pts[(R - 1) & 1], nullptr, offsets, (uint32_t)n_keys, L, R,
                  chunk_bucket, tile_counter, buckets, pk0, pv0);
    } else {
        NZ_LAUNCH(ctx, k_msm_accum<false>, acc_blocks, ACC_THREADS, 0, d_bases, sorted, offsets, (uint32_t)n_keys, L, 0u,
                  chunk_bucket, tile_counter, buckets, pk0, pv0);
    }
    if (ctx->prof_on) {
        NZ_CUDA(ctx, cudaEventRecord(ctx->prof_ev[ctx->prof_used].second, ctx->stream));
        ctx->prof_used++;
        ctx->prof_modmul += 160.0 * (double)n_sum;
    }
    NZ_LAUNCH(ctx, k_seg_join, div_up((size_t)2 * T1, 128), 128, 0, pk0, pv0, 2 * T1, buckets, pk1);
    {
        uint32_t N = 2 * T1;
        uint32_t *ki = pk1, *ko = pk0;  // the join wrote the surviving keys to pk1; the values stay in pv0
        G1XYZZ *vi = pv0, *vo = pv1;
        const uint32_t SL = 16;  // list entries per thread and level
        while (N > 32) {
            const uint32_t threads = (N + SL - 1) / SL;
            NZ_LAUNCH(ctx, k_seg_level, div_up(threads, 128), 128, 0, ki, vi, N, SL, buckets, ko, vo, 0);
            N = 2 * threads;
            std::swap(ki, ko);
            std::swap(vi, vo);
        }
        NZ_LAUNCH(ctx, k_seg_level, 1, 32, 0, ki, vi, N, N, buckets, ko, vo, 1);
    }

    // 5: bucket reduction, all sets at once
    const G1XYZZ* X = buckets;
    const G1XYZZ* P = nullptr;
    G1XYZZ *xo = rx0, *po = rp0, *xo2 = rx1, *po2 = rp1;
    uint32_t bits = p.c - 1;  // log2 of the per-set length
    int one_based = 1;
    G1XYZZ* tail_out = nullptr;  // set when k_bred_tail produced the per-set results
    while (bits > 0) {
        if (!one_based && bits <= BRED_TAIL_BITS) {  // the remaining levels in one launch, one CTA per set
            tail_out = (G1XYZZ*)ctx->scratch_get("msm_tail_out", (size_t)p.G * sizeof(G1XYZZ));
            if (!tail_out) return ctx->fail(NZCB_E_NOMEM, "msm: out of device memory");
            NZ_LAUNCH(ctx, k_bred_tail, p.G, BRED_TAIL_THREADS, 0, const_cast<G1XYZZ*>(X), const_cast<G1XYZZ*>(P), xo, po, bits, tail_out);
            P = tail_out;
            bits = 0;
            break;
        }
        if (one_based) {  // first level: every bucket, throughput bound -- serial running sums over chunks of 8
            const uint32_t log_S = bits >= 3 ? 3 : bits;
            bits -= log_S;
            const uint32_t total = p.G << bits;
            NZ_LAUNCH(ctx, k_bred, div_up(total, 128), 128, 0, X, P, total, log_S, one_based, xo, po);
            one_based = 0;
        } else {          // the rest: latency bound -- halve per launch
            bits -= 1;
            const uint32_t pairs = p.G << bits;
            const uint32_t rb = div_up(pairs, 128);
            NZ_LAUNCH(ctx, k_bred_pair, 2 * rb, 128, 0, X, P, pairs, rb, xo, po);
        }
        X = xo;
        P = po;
        std::swap(xo, xo2);
        std::swap(po, po2);
    }
    // now P[g] = set g's weighted sum (c == 1 cannot happen: c >= 2)
    if (p.unified) {
        NZ_LAUNCH(ctx, k_copy_pts, 1, 32, 0, P, d_out, (uint32_t)K);
    } else {
        NZ_LAUNCH(ctx, k_msm_horner, (unsigned)K, 32, 0, P, p.W, p.c, d_out);
    }
    return 0;
}

// ---- fixed-base tables -----------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_table_build(const G1Affine* __restrict__ bases, uint32_t n, uint32_t stride,
                                                     uint32_t c, uint32_t W, G1Affine* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const G1Affine P = bases[i];
    out[i] = P;
    G1XYZZ acc = G1XYZZ::from_affine(P);
    for (uint32_t w = 1; w < W; w++) {
        for (uint32_t j = 0; j < c; j++) acc = acc.dbl();
        out[(size_t)w * stride + i] = acc.to_affine();
    }
}

int g1_table_build(nzcb_ctx* ctx, const G1Affine* d_bases, size_t n, G1Table* out, uint32_t window) {
    if (n == 0 || n >= ((size_t)1 << 26)) return ctx->fail(NZCB_E_INVALID, "g1 table: bad size %zu", n);
    out->n = n;
    out->stride = n;
    out->c = window ? window : msm_table_window(n);
    out->W = 254 / out->c + 1;
    if ((size_t)out->W * n >= ((size_t)1 << 31)) return ctx->fail(NZCB_E_INVALID, "g1 table: too many points");
    if (cudaMalloc(&out->pts, (size_t)out->W * n * sizeof(G1Affine)) != cudaSuccess) {
        cudaGetLastError();
        out->pts = nullptr;
        return ctx->fail(NZCB_E_NOMEM, "g1 table: cannot allocate %zu bytes", (size_t)out->W * n * sizeof(G1Affine));
    }
    NZ_LAUNCH(ctx, k_table_build, div_up(n, 128), 128, 0, d_bases, (uint32_t)n, (uint32_t)n, out->c, out->W, out->pts);
    return 0;
}

void g1_table_free(G1Table* t) {
    if (t && t->pts) cudaFree(t->pts);
    if (t) t->pts = nullptr;
}

int msm_table_dev(nzcb_ctx* ctx, const G1Table& tab, const uint32_t* const* d_scalars, const size_t* n, int K,
                  bool scalars_mont, G1XYZZ* d_out, bool sparse) {
    if (K < 1 || K > NZ_MSM_MAXJOBS) return ctx->fail(NZCB_E_INVALID, "msm: 1..%d jobs per batch", NZ_MSM_MAXJOBS);
    MsmJob jobs[NZ_MSM_MAXJOBS];
    for (int k = 0; k < K; k++) {
        if (n[k] > tab.n) return ctx->fail(NZCB_E_INVALID, "msm: %zu scalars for a table of %zu bases", n[k], tab.n);
        const nzcb_ctx* root = ctx->root();
        size_t lo = 0, hi = n[k];
        if (root->split_world > 1) {  // latency mode: contiguous slice of the point range
            lo = n[k] * (size_t)root->split_rank / (size_t)root->split_world;
            hi = n[k] * (size_t)(root->split_rank + 1) / (size_t)root->split_world;
        }
        jobs[k] = MsmJob{d_scalars[k], lo, hi, scalars_mont};
    }
    MsmPlan p;
    p.c = tab.c; p.W = tab.W; p.nbw = 1u << (tab.c - 1); p.G = (uint32_t)K; p.unified = true; p.stride = (uint32_t)tab.stride; p.sparse = sparse;
    return msm_run(ctx, tab.pts, p, jobs, K, d_out);
}

int msm_dev(nzcb_ctx* ctx, const G1Affine* d_bases, const uint32_t* d_scalars, size_t n, bool mont, G1XYZZ* d_out) {
    const MsmJob job{d_scalars, 0, n, mont};
    const MsmPlan p = make_plan_window(n, 1);
    return msm_run(ctx, d_bases, p, &job, 1, d_out);
}

int msm_to_host_affine(nzcb_ctx* ctx, const G1XYZZ* d_pt, G1Affine* h_out, int count) {
    G1XYZZ h[NZ_MSM_MAXJOBS];
    if (count < 1 || count > NZ_MSM_MAXJOBS) return ctx->fail(NZCB_E_INVALID, "msm: bad point count");
    NZ_CUDA(ctx, cudaMemcpyAsync(h, d_pt, (size_t)count * sizeof(G1XYZZ), cudaMemcpyDeviceToHost, ctx->stream));
    NZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < count; i++) h_out[i] = h[i].to_affine();  // one Fq inversion each on the host: O(1) finishing step
    return 0;
}

// ---- latency mode: exchange of the partial sums over NCCL, on the device ------------------------------------
// libnccl is NOT a link-time dependency: the caller names the library (the one torch bundles) and it is dlopen'ed.
namespace {
struct Id128 {  // ncclUniqueId: 128 opaque bytes, passed by value
    char b[128];
};
struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(void*) = nullptr;
    int (*CommInitRank)(void**, int, Id128, int) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
NcclApi g_nccl;
std::mutex g_nccl_mu;
int nccl_load(const char* path, char* err, size_t err_len) {
    std::lock_guard<std::mutex> g(g_nccl_mu);
    if (g_nccl.lib) return 0;
    void* h = dlopen(path && path[0] ? path : "libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) {
        snprintf(err, err_len, "cannot load NCCL (%s): %s", path ? path : "libnccl.so.2", dlerror());
        return NZCB_E_INVALID;
    }
    NcclApi a;
    a.lib = h;
    a.GetUniqueId = (int (*)(void*))dlsym(h, "ncclGetUniqueId");
    a.CommInitRank = (int (*)(void**, int, Id128, int))dlsym(h, "ncclCommInitRank");
    a.AllGather = (int (*)(const void*, void*, size_t, int, void*, cudaStream_t))dlsym(h, "ncclAllGather");
    a.CommDestroy = (int (*)(void*))dlsym(h, "ncclCommDestroy");
    a.GetErrorString = (const char* (*)(int))dlsym(h, "ncclGetErrorString");
    if (!a.GetUniqueId || !a.CommInitRank || !a.AllGather || !a.CommDestroy || !a.GetErrorString) {
        snprintf(err, err_len, "NCCL library lacks a required symbol");
        return NZCB_E_INVALID;
    }
    g_nccl = a;
    return 0;
}
// out[k] = sum over ranks (in rank order) of all[r * count + k]: every rank adds the same points in the same order
__global__ void k_sum_partials(const G1XYZZ* __restrict__ all, int world, int count, G1XYZZ* __restrict__ out) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    G1XYZZ acc = G1XYZZ::inf();
    for (int r = 0; r < world; r++) acc.add(all[(size_t)r * count + k]);
    out[k] = acc;
}
}  // namespace

// called when a context goes away: the latency mode's communicator is the library's own
void msm_split_release(nzcb_ctx* ctx) {
    if (ctx && ctx->split_nccl_comm && g_nccl.CommDestroy) g_nccl.CommDestroy(ctx->split_nccl_comm);
    if (ctx) ctx->split_nccl_comm = nullptr;
}

int msm_table_finish(nzcb_ctx* ctx, const G1XYZZ* d_pt, G1Affine* h_out, int count) {
    nzcb_ctx* root = ctx->root();
    if (root->split_world <= 1) return msm_to_host_affine(ctx, d_pt, h_out, count);
    if (count < 1 || count > NZ_MSM_MAXJOBS) return ctx->fail(NZCB_E_INVALID, "msm: bad point count");
    if (root->split_nccl_comm) {  // device-side exchange: all-gather on this stream, sum by a kernel, one read-back
        G1XYZZ* d_all = (G1XYZZ*)ctx->scratch_get("msm_split_all", (size_t)root->split_world * NZ_MSM_MAXJOBS * sizeof(G1XYZZ));
        G1XYZZ* d_sum = (G1XYZZ*)ctx->scratch_get("msm_split_sum", NZ_MSM_MAXJOBS * sizeof(G1XYZZ));
        if (!d_all || !d_sum) return ctx->fail(NZCB_E_NOMEM, "msm split: out of device memory");
        const int rc = g_nccl.AllGather(d_pt, d_all, (size_t)count * sizeof(G1XYZZ), /* ncclChar */ 0, root->split_nccl_comm, ctx->stream);
        if (rc != 0) return ctx->fail(NZCB_E_CUDA, "msm split: ncclAllGather failed: %s", g_nccl.GetErrorString(rc));
        NZ_LAUNCH(ctx, k_sum_partials, 1, 32, 0, d_all, root->split_world, count, d_sum);
        return msm_to_host_affine(ctx, d_sum, h_out, count);
    }
    if (!root->split_allgather) return ctx->fail(NZCB_E_INVALID, "msm split: no exchange function set");
    G1XYZZ mine[NZ_MSM_MAXJOBS];
    NZ_CUDA(ctx, cudaMemcpyAsync(mine, d_pt, (size_t)count * sizeof(G1XYZZ), cudaMemcpyDeviceToHost, ctx->stream));
    NZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    std::vector<G1XYZZ> all((size_t)root->split_world * count);
    if (root->split_allgather(root->split_user, mine, all.data(), (size_t)count * sizeof(G1XYZZ)) != 0)
        return ctx->fail(NZCB_E_CUDA, "msm split: the partial-sum exchange failed");
    for (int k = 0; k < count; k++) {  // rank order is fixed, so every rank adds the same points in the same order
        G1XYZZ acc = G1XYZZ::inf();
        for (int r = 0; r < root->split_world; r++) acc.add(all[(size_t)r * count + k]);
        h_out[k] = acc.to_affine();
    }
    return 0;
}

}  // namespace nzcb

using namespace nzcb;

extern "C" int32_t nzcb_nccl_unique_id(const char* libnccl_path, uint8_t id[128]) {
    char err[256];
    if (!id || nccl_load(libnccl_path, err, sizeof err) != 0) return NZCB_E_INVALID;
    return g_nccl.GetUniqueId(id) == 0 ? 0 : NZCB_E_CUDA;
}

extern "C" int32_t nzcb_ctx_set_msm_split_nccl(nzcb_ctx* ctx, int32_t rank, int32_t world, const char* libnccl_path,
                                               const uint8_t id[128]) {
    if (!ctx || ctx->parent || world < 1 || rank < 0 || rank >= world) return NZCB_E_INVALID;
    NZ_CUDA(ctx, cudaSetDevice(ctx->device));
    if (ctx->split_nccl_comm) {
        cudaStreamSynchronize(ctx->stream);
        for (nzcb_ctx* l : ctx->lanes) cudaStreamSynchronize(l->stream);
        g_nccl.CommDestroy(ctx->split_nccl_comm);
        ctx->split_nccl_comm = nullptr;
    }
    ctx->split_rank = 0;
    ctx->split_world = 1;
    ctx->split_allgather = nullptr;
    if (world == 1) return 0;
    if (!id) return NZCB_E_INVALID;
    if (nccl_load(libnccl_path, ctx->err, sizeof ctx->err) != 0) return NZCB_E_INVALID;
    Id128 uid;
    memcpy(uid.b, id, 128);
    void* comm = nullptr;
    const int rc = g_nccl.CommInitRank(&comm, world, uid, rank);
    if (rc != 0) return ctx->fail(NZCB_E_CUDA, "ncclCommInitRank failed: %s", g_nccl.GetErrorString(rc));
    ctx->split_nccl_comm = comm;
    ctx->split_rank = rank;
    ctx->split_world = world;
    return 0;
}

extern "C" int32_t nzcb_ctx_set_msm_split(nzcb_ctx* ctx, int32_t rank, int32_t world,
                                          int (*allgather)(void*, const void*, void*, size_t), void* user) {
    if (!ctx || ctx->parent || world < 1 || rank < 0 || rank >= world || (world > 1 && !allgather)) return NZCB_E_INVALID;
    if (ctx->split_nccl_comm) return ctx->fail(NZCB_E_INVALID, "msm split: switch the NCCL mode off first");
    ctx->split_rank = rank;
    ctx->split_world = world;
    ctx->split_allgather = allgather;
    ctx->split_user = user;
    return 0;
}

struct nzcb_g1_table {
    nzcb_ctx* ctx;
    G1Table tab;
};

extern "C" int32_t nzcb_g1_table_create(nzcb_ctx* ctx, const uint8_t* bases, size_t n, nzcb_g1_table** out) {
    if (!ctx || !bases || !out || n == 0) return NZCB_E_INVALID;
    *out = nullptr;
    NZ_CUDA(ctx, cudaSetDevice(ctx->device));
    G1Affine* d_b = (G1Affine*)ctx->scratch_get("msm_in_bases", n * 64 + 64);
    if (!d_b) return ctx->fail(NZCB_E_NOMEM, "g1 table: cannot allocate input buffer for n=%zu", n);
    NZ_CUDA(ctx, cudaMemcpyAsync(d_b, bases, n * 64, cudaMemcpyHostToDevice, ctx->stream));
    nzcb_g1_table* t = new nzcb_g1_table();
    t->ctx = ctx;
    const int rc = g1_table_build(ctx, d_b, n, &t->tab);
    if (rc != 0 || cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
        g1_table_free(&t->tab);
        delete t;
        return rc ? rc : ctx->fail(NZCB_E_CUDA, "g1 table: build failed");
    }
    *out = t;
    return 0;
}

extern "C" void nzcb_g1_table_free(nzcb_g1_table* t) {
    if (!t) return;
    if (t->ctx) {
        cudaSetDevice(t->ctx->device);
        cudaStreamSynchronize(t->ctx->stream);
    }
    g1_table_free(&t->tab);
    delete t;
}

// K MSMs over the first n[k] bases of the table, scalars device resident (n x 32 B LE canonical each)
extern "C" int32_t nzcb_msm_g1_table_dev(nzcb_ctx* ctx, const nzcb_g1_table* t, const void* const* d_scalars,
                                         const size_t* n, int32_t K, uint8_t* out /* K x 64 */) {
    if (!ctx || !t || !d_scalars || !n || !out || t->ctx != ctx) return NZCB_E_INVALID;
    G1XYZZ* d_out = (G1XYZZ*)ctx->scratch_get("msm_out", NZ_MSM_MAXJOBS * sizeof(G1XYZZ));
    if (!d_out) return ctx->fail(NZCB_E_NOMEM, "msm: out of device memory");
    NZ_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
    NZ_TRY(msm_table_dev(ctx, t->tab, (const uint32_t* const*)d_scalars, n, K, false, d_out));
    NZ_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
    G1Affine a[NZ_MSM_MAXJOBS];
    NZ_TRY(msm_table_finish(ctx, d_out, a, K));
    cudaEventElapsedTime(&ctx->last_ms, ctx->ev0, ctx->ev1);
    memcpy(out, a, (size_t)K * 64);
    return 0;
}

extern "C" int32_t nzcb_msm_g1_table(nzcb_ctx* ctx, const nzcb_g1_table* t, const uint8_t* scalars, size_t n,
                                     uint8_t out[64]) {
    if (!ctx || !t || !out || (n && !scalars) || t->ctx != ctx) return NZCB_E_INVALID;
    uint32_t* d_s = (uint32_t*)ctx->scratch_get("msm_in_scalars", n * 32 + 32);
    if (!d_s) return ctx->fail(NZCB_E_NOMEM, "msm: cannot allocate input buffers for n=%zu", n);
    if (n) NZ_CUDA(ctx, cudaMemcpyAsync(d_s, scalars, n * 32, cudaMemcpyHostToDevice, ctx->stream));
    const void* sp = d_s;
    return nzcb_msm_g1_table_dev(ctx, t, &sp, &n, 1, out);
}

extern "C" int32_t nzcb_msm_g1_dev(nzcb_ctx* ctx, const void* d_bases, const void* d_scalars, size_t n,
                                   uint8_t out[64]) {
    if (!ctx || !out || (n && (!d_bases || !d_scalars))) return NZCB_E_INVALID;
    G1XYZZ* d_out = (G1XYZZ*)ctx->scratch_get("msm_out", NZ_MSM_MAXJOBS * sizeof(G1XYZZ));
    if (!d_out) return ctx->fail(NZCB_E_NOMEM, "msm: out of device memory");
    NZ_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
    NZ_TRY(msm_dev(ctx, (const G1Affine*)d_bases, (const uint32_t*)d_scalars, n, false, d_out));
    NZ_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
    G1Affine a;
    NZ_TRY(msm_to_host_affine(ctx, d_out, &a, 1));
    cudaEventElapsedTime(&ctx->last_ms, ctx->ev0, ctx->ev1);
    memcpy(out, &a, 64);
    return 0;
}

extern "C" int32_t nzcb_msm_g1(nzcb_ctx* ctx, const uint8_t* bases, const uint8_t* scalars, size_t n, uint8_t out[64]) {
    if (!ctx || !out || (n && (!bases || !scalars))) return NZCB_E_INVALID;
    G1Affine* d_b = (G1Affine*)ctx->scratch_get("msm_in_bases", n * 64 + 64);
    uint32_t* d_s = (uint32_t*)ctx->scratch_get("msm_in_scalars", n * 32 + 32);
    if (!d_b || !d_s) return ctx->fail(NZCB_E_NOMEM, "msm: cannot allocate input buffers for n=%zu", n);
    if (n) {
        NZ_CUDA(ctx, cudaMemcpyAsync(d_b, bases, n * 64, cudaMemcpyHostToDevice, ctx->stream));
        NZ_CUDA(ctx, cudaMemcpyAsync(d_s, scalars, n * 32, cudaMemcpyHostToDevice, ctx->stream));
    }
    return nzcb_msm_g1_dev(ctx, d_b, d_s, n, out);
}
