// BN254 G1 multi-scalar multiplication -- replaces ffjavascript
// G1.multiExpAffine (un-vendored, /root/reference/yarn.lock:3905; called nine
// times per proof by snarkjs plonk.prove).
//
// Pippenger with signed c-bit digits (buckets 1..2^(c-1), negative digits add
// the negated base):
//   1. k_msm_count    one thread per scalar: recode, histogram the (window,|digit|) keys
//   2. k_scan_excl    bucket offsets
//   3. k_msm_scatter  counting-sort the point indices by bucket (order inside a
//                     bucket is arbitrary -- the group sum is the same element)
//   4. k_msm_accum    one thread per bucket: XYZZ += affine over its run (8M+2S each)
//   5. k_msm_reduce1  per (window, chunk of S buckets): running-sum trick
//      k_msm_reduce2  per window: block tree reductions combine the chunks
//      k_msm_final    Horner over windows
// Algorithmic work: n * windows * 10 modmul in step 4.
#include "common.cuh"
#include <stdlib.h>

namespace nzcb {

struct MsmPlan {
    uint32_t c;        // window bits
    uint32_t W;        // number of windows
    uint32_t nbw;      // buckets per window = 2^(c-1)
    uint32_t C;        // chunks per window in the reduction
    uint32_t S;        // buckets per chunk
    uint32_t log_S;
    uint32_t log_C;
    size_t nb;         // total buckets
};

static MsmPlan make_plan(size_t n) {
    uint32_t lg = 0;
    while (((size_t)2 << lg) <= n) lg++;
    int c = (int)lg - 5;
    if (c < 4) c = 4;
    if (c > 16) c = 16;
    const char* env = getenv("NZCB_MSM_WINDOW");
    if (env) {
        int v = atoi(env);
        if (v >= 2 && v <= 20) c = v;
    }
    MsmPlan p;
    p.c = (uint32_t)c;
    p.W = 254 / p.c + 1;
    p.nbw = 1u << (p.c - 1);
    p.C = p.nbw < 256 ? p.nbw : 256;
    p.S = p.nbw / p.C;
    p.log_S = 0;
    while ((1u << p.log_S) < p.S) p.log_S++;
    p.log_C = 0;
    while ((1u << p.log_C) < p.C) p.log_C++;
    p.nb = (size_t)p.W * p.nbw;
    return p;
}

__device__ __forceinline__ uint32_t get_bits(const uint32_t* s, uint32_t off, uint32_t c) {
    const uint32_t limb = off >> 5, sh = off & 31;
    if (limb >= 8) return 0;
    uint32_t v = s[limb] >> sh;
    if (sh + c > 32 && limb + 1 < 8) v |= s[limb + 1] << (32 - sh);
    return v & ((1u << c) - 1);
}

// Signed-digit recode of one scalar; calls f(window, bucket_index, negative) per non-zero digit.
template <class F>
__device__ __forceinline__ void for_each_digit(const uint32_t* __restrict__ scalars, size_t i, bool mont, uint32_t c,
                                               uint32_t W, F f) {
    Fr s = reinterpret_cast<const Fr*>(scalars)[i];
    if (mont) s = s.from_mont();
    uint32_t carry = 0;
    const uint32_t half = 1u << (c - 1);
    for (uint32_t w = 0; w < W; w++) {
        uint32_t d = get_bits(s.v, w * c, c) + carry;
        if (d > half) {
            carry = 1;
            const uint32_t mag = (1u << c) - d;  // |d - 2^c|; 0 when d == 2^c (digit 0, carry 1)
            if (mag) f(w, mag - 1, true);
        } else {
            carry = 0;
            if (d) f(w, d - 1, false);
        }
    }
}

__global__ void k_msm_count(const uint32_t* __restrict__ scalars, size_t n, int mont, uint32_t c, uint32_t W,
                            uint32_t nbw, uint32_t* __restrict__ counts) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    for_each_digit(scalars, i, mont != 0, c, W,
                   [&](uint32_t w, uint32_t b, bool) { atomicAdd(&counts[(size_t)w * nbw + b], 1u); });
}

// exclusive scan of `n` counts into offsets[0..n] (offsets[n] = total); single block
__global__ void __launch_bounds__(1024) k_scan_excl(const uint32_t* __restrict__ counts, uint32_t* __restrict__ offsets,
                                                    size_t n) {
    __shared__ uint32_t part[1024];
    const uint32_t t = threadIdx.x;
    const size_t per = (n + 1023) / 1024;
    const size_t lo = (size_t)t * per;
    const size_t hi = lo + per < n ? lo + per : n;
    uint32_t sum = 0;
    for (size_t k = lo; k < hi; k++) sum += counts[k];
    part[t] = sum;
    __syncthreads();
    for (uint32_t off = 1; off < 1024; off <<= 1) {
        uint32_t v = t >= off ? part[t - off] : 0;
        __syncthreads();
        part[t] += v;
        __syncthreads();
    }
    uint32_t run = part[t] - sum;
    for (size_t k = lo; k < hi; k++) {
        offsets[k] = run;
        run += counts[k];
    }
    if (t == 1023) offsets[n] = part[1023];
}

__global__ void k_msm_scatter(const uint32_t* __restrict__ scalars, size_t n, int mont, uint32_t c, uint32_t W,
                              uint32_t nbw, const uint32_t* __restrict__ offsets, uint32_t* __restrict__ cursor,
                              uint32_t* __restrict__ sorted) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    for_each_digit(scalars, i, mont != 0, c, W, [&](uint32_t w, uint32_t b, bool neg) {
        const size_t key = (size_t)w * nbw + b;
        const uint32_t pos = atomicAdd(&cursor[key], 1u);
        sorted[(size_t)offsets[key] + pos] = (uint32_t)i | (neg ? 0x80000000u : 0u);
    });
}

__global__ void __launch_bounds__(128) k_msm_accum(const G1Affine* __restrict__ bases, const uint32_t* __restrict__ sorted,
                                                   const uint32_t* __restrict__ offsets, size_t nb,
                                                   G1XYZZ* __restrict__ buckets) {
    const size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    const uint32_t lo = offsets[b], hi = offsets[b + 1];
    G1XYZZ acc = G1XYZZ::inf();
    for (uint32_t k = lo; k < hi; k++) {
        const uint32_t e = sorted[k];
        G1Affine p = bases[e & 0x7fffffffu];
        if (e & 0x80000000u) p.y = p.y.neg();  // (0,0) stays (0,0)
        acc.add_affine(p);
    }
    buckets[b] = acc;
}

// per (window, chunk): run = sum B_k, acc = sum (k+1) * B_k over the chunk's S buckets
__global__ void __launch_bounds__(128) k_msm_reduce1(const G1XYZZ* __restrict__ buckets, uint32_t W, uint32_t nbw,
                                                     uint32_t C, uint32_t S, G1XYZZ* __restrict__ run_out,
                                                     G1XYZZ* __restrict__ acc_out) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)W * C) return;
    const uint32_t w = (uint32_t)(t / C), ch = (uint32_t)(t % C);
    const G1XYZZ* B = buckets + (size_t)w * nbw + (size_t)ch * S;
    G1XYZZ run = G1XYZZ::inf(), acc = G1XYZZ::inf();
    for (int k = (int)S - 1; k >= 0; k--) {
        run.add(B[k]);
        acc.add(run);
    }
    run_out[t] = run;
    acc_out[t] = acc;
}

__device__ __forceinline__ G1XYZZ block_tree_sum(G1XYZZ* sh, G1XYZZ v, uint32_t C) {
    const uint32_t t = threadIdx.x;
    __syncthreads();
    sh[t] = v;
    __syncthreads();
    for (uint32_t off = C >> 1; off >= 1; off >>= 1) {
        if (t < off) {
            G1XYZZ a = sh[t];
            a.add(sh[t + off]);
            sh[t] = a;
        }
        __syncthreads();
    }
    return sh[0];
}

// one block (C threads) per window: total_w = sum_ch acc_ch + S * sum_ch ch * run_ch
__global__ void k_msm_reduce2(const G1XYZZ* __restrict__ run_in, const G1XYZZ* __restrict__ acc_in, uint32_t C,
                              uint32_t log_C, uint32_t log_S, G1XYZZ* __restrict__ win_out) {
    extern __shared__ __align__(32) unsigned char smem_raw[];
    G1XYZZ* sh = reinterpret_cast<G1XYZZ*>(smem_raw);
    const uint32_t w = blockIdx.x, t = threadIdx.x;
    const G1XYZZ my_run = run_in[(size_t)w * C + t];
    G1XYZZ total = block_tree_sum(sh, acc_in[(size_t)w * C + t], C);
    G1XYZZ T = G1XYZZ::inf();
    for (int j = (int)log_C - 1; j >= 0; j--) {
        G1XYZZ pj = block_tree_sum(sh, ((t >> j) & 1) ? my_run : G1XYZZ::inf(), C);
        if (t == 0) {
            T = T.dbl();
            T.add(pj);
        }
    }
    if (t == 0) {
        for (uint32_t i = 0; i < log_S; i++) T = T.dbl();
        total.add(T);
        win_out[w] = total;
    }
}

__global__ void k_msm_final(const G1XYZZ* __restrict__ win, uint32_t W, uint32_t c, G1XYZZ* __restrict__ out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    G1XYZZ r = G1XYZZ::inf();
    for (int w = (int)W - 1; w >= 0; w--) {
        for (uint32_t i = 0; i < c; i++) r = r.dbl();
        r.add(win[w]);
    }
    *out = r;
}

__global__ void k_set_inf(G1XYZZ* out) { *out = G1XYZZ::inf(); }

int msm_dev(nzcb_ctx* ctx, const G1Affine* d_bases, const uint32_t* d_scalars, size_t n, bool mont, G1XYZZ* d_out) {
    if (n == 0) {
        NZ_LAUNCH(ctx, k_set_inf, 1, 1, 0, d_out);
        return 0;
    }
    if (n >= ((size_t)1 << 31)) return ctx->fail(NZCB_E_INVALID, "msm: n too large");
    const MsmPlan p = make_plan(n);
    uint32_t* counts = (uint32_t*)ctx->scratch_get("msm_counts", (p.nb + 1) * 4);
    uint32_t* offsets = (uint32_t*)ctx->scratch_get("msm_offsets", (p.nb + 1) * 4);
    uint32_t* cursor = (uint32_t*)ctx->scratch_get("msm_cursor", (p.nb + 1) * 4);
    uint32_t* sorted = (uint32_t*)ctx->scratch_get("msm_sorted", n * p.W * 4);
    G1XYZZ* buckets = (G1XYZZ*)ctx->scratch_get("msm_buckets", p.nb * sizeof(G1XYZZ));
    G1XYZZ* run = (G1XYZZ*)ctx->scratch_get("msm_run", (size_t)p.W * p.C * sizeof(G1XYZZ));
    G1XYZZ* acc = (G1XYZZ*)ctx->scratch_get("msm_acc", (size_t)p.W * p.C * sizeof(G1XYZZ));
    G1XYZZ* win = (G1XYZZ*)ctx->scratch_get("msm_win", (size_t)p.W * sizeof(G1XYZZ));
    if (!counts || !offsets || !cursor || !sorted || !buckets || !run || !acc || !win)
        return ctx->fail(NZCB_E_NOMEM, "msm: cannot allocate workspace for n=%zu", n);
    NZ_CUDA(ctx, cudaMemsetAsync(counts, 0, (p.nb + 1) * 4, ctx->stream));
    NZ_CUDA(ctx, cudaMemsetAsync(cursor, 0, (p.nb + 1) * 4, ctx->stream));
    NZ_LAUNCH(ctx, k_msm_count, div_up(n, 256), 256, 0, d_scalars, n, mont ? 1 : 0, p.c, p.W, p.nbw, counts);
    NZ_LAUNCH(ctx, k_scan_excl, 1, 1024, 0, counts, offsets, p.nb);
    NZ_LAUNCH(ctx, k_msm_scatter, div_up(n, 256), 256, 0, d_scalars, n, mont ? 1 : 0, p.c, p.W, p.nbw, offsets, cursor,
              sorted);
    if (ctx->prof_on) {
        if (ctx->prof_used == ctx->prof_ev.size()) {
            cudaEvent_t a, b;
            NZ_CUDA(ctx, cudaEventCreate(&a));
            NZ_CUDA(ctx, cudaEventCreate(&b));
            ctx->prof_ev.push_back({a, b});
        }
        NZ_CUDA(ctx, cudaEventRecord(ctx->prof_ev[ctx->prof_used].first, ctx->stream));
    }
    NZ_LAUNCH(ctx, k_msm_accum, div_up(p.nb, 128), 128, 0, d_bases, sorted, offsets, p.nb, buckets);
    if (ctx->prof_on) {
        NZ_CUDA(ctx, cudaEventRecord(ctx->prof_ev[ctx->prof_used].second, ctx->stream));
        ctx->prof_used++;
        ctx->prof_modmul += 160.0 * (double)n;
    }
    NZ_LAUNCH(ctx, k_msm_reduce1, div_up((size_t)p.W * p.C, 128), 128, 0, buckets, p.W, p.nbw, p.C, p.S, run, acc);
    NZ_LAUNCH(ctx, k_msm_reduce2, p.W, p.C, p.C * sizeof(G1XYZZ), run, acc, p.C, p.log_C, p.log_S, win);
    NZ_LAUNCH(ctx, k_msm_final, 1, 32, 0, win, p.W, p.c, d_out);
    return 0;
}

int msm_to_host_affine(nzcb_ctx* ctx, const G1XYZZ* d_pt, G1Affine* h_out) {
    G1XYZZ h;
    NZ_CUDA(ctx, cudaMemcpyAsync(&h, d_pt, sizeof(G1XYZZ), cudaMemcpyDeviceToHost, ctx->stream));
    NZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *h_out = h.to_affine();  // one Fq inversion on the host: O(1) finishing step
    return 0;
}

}  // namespace nzcb

using namespace nzcb;

extern "C" int32_t nzcb_msm_g1_dev(nzcb_ctx* ctx, const void* d_bases, const void* d_scalars, size_t n,
                                   uint8_t out[64]) {
    if (!ctx || !out || (n && (!d_bases || !d_scalars))) return NZCB_E_INVALID;
    G1XYZZ* d_out = (G1XYZZ*)ctx->scratch_get("msm_out", sizeof(G1XYZZ));
    if (!d_out) return ctx->fail(NZCB_E_NOMEM, "msm: out of device memory");
    NZ_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
    NZ_TRY(msm_dev(ctx, (const G1Affine*)d_bases, (const uint32_t*)d_scalars, n, false, d_out));
    NZ_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
    G1Affine a;
    NZ_TRY(msm_to_host_affine(ctx, d_out, &a));
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
