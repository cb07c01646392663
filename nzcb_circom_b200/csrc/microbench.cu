// Integer-pipe microbenchmarks: the measured denominators of the IMAD roofline
// ("the 64/clk/SM figure must be confirmed by an IMAD
// microbenchmark").  kind 0 = IMAD (32x32+32 lo), 1 = IMAD.WIDE.U32 (32x32+64),
// 2 = Fr Montgomery multiply, 3 = Fq Montgomery multiply.
#include "common.cuh"
#include <type_traits>

namespace nzcb {

template <int KIND>
__global__ void __launch_bounds__(256) k_microbench(uint32_t* out, uint32_t iters, uint32_t seed) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    if (KIND == 0) {
        uint32_t a0 = tid * 2654435761u + seed, a1 = a0 ^ 0x9e3779b9u, a2 = a0 + 77u, a3 = a1 + 1234567u;
        uint32_t a4 = a0 * 3u, a5 = a1 * 5u, a6 = a2 * 7u, a7 = a3 * 11u;
        const uint32_t m = seed | 1u;
        for (uint32_t i = 0; i < iters; i++) {
#pragma unroll
            for (int u = 0; u < 8; u++) {  // 8 independent chains x 8 = 64 IMAD per iteration
                a0 = a0 * m + a1; a1 = a1 * m + a2; a2 = a2 * m + a3; a3 = a3 * m + a4;
                a4 = a4 * m + a5; a5 = a5 * m + a6; a6 = a6 * m + a7; a7 = a7 * m + a0;
            }
        }
        out[tid] = a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7;
    } else if (KIND == 1) {
        uint64_t a0 = tid * 2654435761ull + seed, a1 = a0 ^ 0x9e3779b97f4a7c15ull, a2 = a0 + 77u, a3 = a1 + 1234567u;
        uint64_t a4 = a0 * 3u, a5 = a1 * 5u, a6 = a2 * 7u, a7 = a3 * 11u;
        const uint32_t m = seed | 1u;
        for (uint32_t i = 0; i < iters; i++) {
#pragma unroll
            for (int u = 0; u < 8; u++) {  // 64 IMAD.WIDE.U32 per iteration
                a0 = (uint64_t)(uint32_t)a1 * m + a0; a1 = (uint64_t)(uint32_t)a2 * m + a1;
                a2 = (uint64_t)(uint32_t)a3 * m + a2; a3 = (uint64_t)(uint32_t)a4 * m + a3;
                a4 = (uint64_t)(uint32_t)a5 * m + a4; a5 = (uint64_t)(uint32_t)a6 * m + a5;
                a6 = (uint64_t)(uint32_t)a7 * m + a6; a7 = (uint64_t)(uint32_t)a0 * m + a7;
            }
        }
        uint64_t x = a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7;
        out[tid] = (uint32_t)x ^ (uint32_t)(x >> 32);
    } else if (KIND == 4) {
        uint32_t a0 = tid * 2654435761u + seed, a1 = a0 ^ 0x9e3779b9u, a2 = a0 + 77u, a3 = a1 + 1234567u;
        uint32_t a4 = a0 * 3u, a5 = a1 * 5u, a6 = a2 * 7u, a7 = a3 * 11u;
        const uint32_t m = seed | 0x80000001u;
        for (uint32_t i = 0; i < iters; i++) {
#pragma unroll
            for (int u = 0; u < 8; u++) {  // 64 IMAD.HI.U32 per iteration
                a0 = __umulhi(a0, m) + a1; a1 = __umulhi(a1, m) + a2; a2 = __umulhi(a2, m) + a3; a3 = __umulhi(a3, m) + a4;
                a4 = __umulhi(a4, m) + a5; a5 = __umulhi(a5, m) + a6; a6 = __umulhi(a6, m) + a7; a7 = __umulhi(a7, m) + a0;
            }
        }
        out[tid] = a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7;
    } else if (KIND == 5) {
        Fr a = Fr::from_u64(tid + 3), b = Fr::from_u64(seed + 5), c = Fr::from_u64(tid * 7 + 1), d = Fr::from_u64(seed * 3 + 11);
        for (uint32_t i = 0; i < iters; i++) {
            a = Fr::mul_portable(a, b);
            c = Fr::mul_portable(c, d);
            b = Fr::mul_portable(b, a);
            d = Fr::mul_portable(d, c);
        }
        Fr r = a + b + c + d;
        uint32_t x = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) x ^= r.v[k];
        out[tid] = x;
    } else {
        typedef typename std::conditional<KIND == 2, Fr, Fq>::type F;
        F a = F::from_u64(tid + 3), b = F::from_u64(seed + 5), c = F::from_u64(tid * 7 + 1), d = F::from_u64(seed * 3 + 11);
        for (uint32_t i = 0; i < iters; i++) {  // 4 multiplies per iteration, two independent chains
            a = a * b;
            c = c * d;
            b = b * a;
            d = d * c;
        }
        F r = a + b + c + d;
        uint32_t x = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) x ^= r.v[k];
        out[tid] = x;
    }
}


// ---- mixed-addition loop variants (the body of k_msm_accum): what limits it on the integer pipe? ------------
struct MulInline {
    static __device__ __forceinline__ Fq mul(const Fq& a, const Fq& b) { return a * b; }
};
__device__ __noinline__ Fq fq_mul_call(Fq a, Fq b) { return a * b; }
struct MulRolled {
    static __device__ __forceinline__ Fq mul(const Fq& a, const Fq& b) { return Fq::mul_rolled(a, b); }
};
struct MulCall {
    static __device__ __forceinline__ Fq mul(const Fq& a, const Fq& b) { return fq_mul_call(a, b); }
};
template <class M>
__device__ __forceinline__ void madd_v(G1XYZZ& r, const G1Affine& b) {
    if (b.is_inf()) return;
    if (r.is_inf()) {
        r = G1XYZZ::from_affine(b);
        return;
    }
    const Fq U2 = M::mul(b.x, r.ZZ), S2 = M::mul(b.y, r.ZZZ);
    const Fq Pp = U2 - r.X, Rr = S2 - r.Y;
    if (Pp.is_zero()) {
        r = Rr.is_zero() ? G1XYZZ::from_affine(b).dbl() : G1XYZZ::inf();
        return;
    }
    const Fq PP = M::mul(Pp, Pp), PPP = M::mul(Pp, PP), Q = M::mul(r.X, PP);
    const Fq X3 = M::mul(Rr, Rr) - PPP - Q.dbl();
    r.Y = M::mul(Rr, Q - X3) - M::mul(r.Y, PPP);
    r.X = X3;
    r.ZZ = M::mul(r.ZZ, PP);
    r.ZZZ = M::mul(r.ZZZ, PPP);
}
template <class M, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) k_madd_bench(const G1Affine* __restrict__ tab, uint32_t tab_mask,
                                                              uint32_t iters, uint32_t* __restrict__ out) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    G1XYZZ acc = G1XYZZ::from_affine(tab[tid & tab_mask]);
    uint32_t idx = tid * 2654435761u;
    G1Affine p = tab[idx & tab_mask];
    for (uint32_t i = 0; i < iters; i++) {
        const G1Affine cur = p;
        idx = idx * 1664525u + 1013904223u;
        p = tab[(idx >> 8) & tab_mask];
        madd_v<M>(acc, cur);
    }
    uint32_t x = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) x ^= acc.X.v[k] ^ acc.Y.v[k] ^ acc.ZZ.v[k] ^ acc.ZZZ.v[k];
    out[tid] = x;
}
__global__ void k_madd_table(G1Affine* tab, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    G1Affine g;
    g.x = Fq::from_u64(1);
    g.y = Fq::from_u64(2);
    tab[i] = g1_mul_small(G1XYZZ::from_affine(g), 3 + i).to_affine();
}

// device self-test: the carry-chain multiply must agree bit-for-bit with the portable CIOS
template <class F>
__global__ void k_selftest_mul(uint32_t n, uint32_t seed, unsigned long long* mismatches) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= n) return;
    F a, b;
    uint32_t x = tid * 747796405u + seed;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        x = x * 1664525u + 1013904223u; a.v[k] = x ^ (x >> 15);
        x = x * 1664525u + 1013904223u; b.v[k] = x ^ (x >> 13);
    }
    a.v[7] &= 0x1fffffffu; b.v[7] &= 0x1fffffffu;  // < 2^253 < modulus
    if (tid == 0) { a = F::zero(); }
    if (tid == 1) { a = F::modulus(); a.v[0] -= 1; b = a; }
    F acc_p = a, acc_q = a;
    unsigned long long bad = 0;
    for (int it = 0; it < 16; it++) {
        F p = F::mul_portable(acc_p, b);
        F q = acc_q * b;
        if (p != q) bad++;
        acc_p = p; acc_q = q;
        b = b + acc_p;
    }
    if (bad) atomicAdd(mismatches, bad);
}
}  // namespace nzcb
using namespace nzcb;

extern "C" int32_t nzcb_selftest_mul(nzcb_ctx* ctx, uint32_t n, uint64_t* mismatches) {
    if (!ctx || !mismatches) return NZCB_E_INVALID;
    unsigned long long* d = (unsigned long long*)ctx->scratch_get("selftest", 8);
    if (!d) return ctx->fail(NZCB_E_NOMEM, "selftest: out of memory");
    NZ_CUDA(ctx, cudaMemsetAsync(d, 0, 8, ctx->stream));
    NZ_LAUNCH(ctx, k_selftest_mul<Fr>, div_up(n, 256), 256, 0, n, 17u, d);
    NZ_LAUNCH(ctx, k_selftest_mul<Fq>, div_up(n, 256), 256, 0, n, 29u, d);
    unsigned long long h = 0;
    NZ_CUDA(ctx, cudaMemcpyAsync(&h, d, 8, cudaMemcpyDeviceToHost, ctx->stream));
    NZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *mismatches = h;
    return 0;
}


// returns operations per second in *ops_per_s (IMAD / IMAD.WIDE / modmul), device time in ctx->last_ms
extern "C" int32_t nzcb_microbench(nzcb_ctx* ctx, int32_t kind, uint32_t iters, uint32_t blocks_per_sm,
                                   double* ops_per_s) {
    if (!ctx || !ops_per_s || kind < 0 || kind > 5) return NZCB_E_INVALID;
    const uint32_t grid = (uint32_t)ctx->sm_count * (blocks_per_sm ? blocks_per_sm : 4);
    uint32_t* d = (uint32_t*)ctx->scratch_get("microbench", (size_t)grid * 256 * 4);
    if (!d) return ctx->fail(NZCB_E_NOMEM, "microbench: out of memory");
    for (int rep = 0; rep < 2; rep++) {  // first repetition is the warm-up
        NZ_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
        switch (kind) {
            case 0: NZ_LAUNCH(ctx, k_microbench<0>, grid, 256, 0, d, iters, 12345u); break;
            case 1: NZ_LAUNCH(ctx, k_microbench<1>, grid, 256, 0, d, iters, 12345u); break;
            case 2: NZ_LAUNCH(ctx, k_microbench<2>, grid, 256, 0, d, iters, 12345u); break;
            case 3: NZ_LAUNCH(ctx, k_microbench<3>, grid, 256, 0, d, iters, 12345u); break;
            case 4: NZ_LAUNCH(ctx, k_microbench<4>, grid, 256, 0, d, iters, 12345u); break;
            default: NZ_LAUNCH(ctx, k_microbench<5>, grid, 256, 0, d, iters, 12345u); break;
        }
        NZ_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
        NZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    cudaEventElapsedTime(&ctx->last_ms, ctx->ev0, ctx->ev1);
    const double per_thread = (kind <= 1 || kind == 4) ? 64.0 * iters : 4.0 * iters;
    *ops_per_s = per_thread * grid * 256.0 / (ctx->last_ms * 1e-3);
    return 0;
}


// mixed-addition loop: variant 0 = inlined multiplies (128 thr, 4 CTA/SM), 1 = multiply as a call, 2 = inlined,
// 128 thr x 3 CTA/SM (168 regs), 3 = inlined 256 thr x 2, 4 = call 256 x 3.  Result: mixed additions per second.
extern "C" int32_t nzcb_microbench_madd(nzcb_ctx* ctx, int32_t variant, uint32_t iters, uint32_t log_table,
                                        double* madds_per_s) {
    if (!ctx || !madds_per_s || variant < 0 || variant > 7 || log_table > 24) return NZCB_E_INVALID;
    const uint32_t n = 1u << log_table;
    G1Affine* tab = (G1Affine*)ctx->scratch_get("madd_tab", (size_t)n * sizeof(G1Affine));
    if (!tab) return ctx->fail(NZCB_E_NOMEM, "microbench: out of memory");
    NZ_LAUNCH(ctx, k_madd_table, div_up(n, 128), 128, 0, tab, n);
    uint32_t threads = 128, bps = 4;
    if (variant == 2) bps = 3;
    if (variant == 3) { threads = 256; bps = 2; }
    if (variant == 4) { threads = 256; bps = 3; }
    if (variant == 5) bps = 5;
    if (variant == 6) bps = 6;
    if (variant == 7) bps = 8;
    const uint32_t grid = (uint32_t)ctx->sm_count * bps;
    uint32_t* d = (uint32_t*)ctx->scratch_get("microbench", (size_t)grid * threads * 4);
    if (!d) return ctx->fail(NZCB_E_NOMEM, "microbench: out of memory");
    for (int rep = 0; rep < 2; rep++) {
        NZ_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
        switch (variant) {
            case 0: NZ_LAUNCH(ctx, (k_madd_bench<MulInline, 128, 4>), grid, threads, 0, tab, n - 1, iters, d); break;
            case 1: NZ_LAUNCH(ctx, (k_madd_bench<MulCall, 128, 4>), grid, threads, 0, tab, n - 1, iters, d); break;
            case 2: NZ_LAUNCH(ctx, (k_madd_bench<MulInline, 128, 3>), grid, threads, 0, tab, n - 1, iters, d); break;
            case 3: NZ_LAUNCH(ctx, (k_madd_bench<MulInline, 256, 2>), grid, threads, 0, tab, n - 1, iters, d); break;
            case 4: NZ_LAUNCH(ctx, (k_madd_bench<MulCall, 256, 3>), grid, threads, 0, tab, n - 1, iters, d); break;
            case 5: NZ_LAUNCH(ctx, (k_madd_bench<MulInline, 128, 5>), grid, threads, 0, tab, n - 1, iters, d); break;
            case 6: NZ_LAUNCH(ctx, (k_madd_bench<MulInline, 128, 6>), grid, threads, 0, tab, n - 1, iters, d); break;
            default: NZ_LAUNCH(ctx, (k_madd_bench<MulInline, 128, 8>), grid, threads, 0, tab, n - 1, iters, d); break;
        }
        NZ_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
        NZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    cudaEventElapsedTime(&ctx->last_ms, ctx->ev0, ctx->ev1);
    *madds_per_s = (double)iters * grid * threads / (ctx->last_ms * 1e-3);
    return 0;
}


// ---- level round trip of the witness interpreter: store -> barrier -> dependent load, through global memory
// (kind 0) or shared memory (kind 1); one CTA of 256 threads, `iters` levels.  Result: nanoseconds per level.
__global__ void __launch_bounds__(256) k_level_roundtrip(uint32_t* __restrict__ buf, uint32_t iters, int kind, uint32_t* out) {
    __shared__ uint32_t sm[256 * 8];
    const uint32_t t = threadIdx.x;
    uint32_t v[8];
#pragma unroll
    for (int k = 0; k < 8; k++) v[k] = t + k;
    for (uint32_t i = 0; i < iters; i++) {
        const uint32_t dst = ((i * 256u + t) & 0xffffu) * 8, src = ((i * 256u + ((t + 37u) & 255u)) & 0xffffu) * 8;
        if (kind == 0) {
#pragma unroll
            for (int k = 0; k < 8; k++) buf[dst + k] = v[k];
        } else {
#pragma unroll
            for (int k = 0; k < 8; k++) sm[k * 256 + t] = v[k];
        }
        __syncthreads();
        if (kind == 0) {
#pragma unroll
            for (int k = 0; k < 8; k++) v[k] += buf[src + k];
        } else {
#pragma unroll
            for (int k = 0; k < 8; k++) v[k] += sm[k * 256 + ((t + 37u) & 255u)];
        }
        __syncthreads();
    }
    uint32_t x = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) x ^= v[k];
    out[t] = x;
}

extern "C" int32_t nzcb_microbench_level(nzcb_ctx* ctx, int32_t kind, uint32_t iters, double* ns_per_level) {
    if (!ctx || !ns_per_level || kind < 0 || kind > 1) return NZCB_E_INVALID;
    uint32_t* buf = (uint32_t*)ctx->scratch_get("level_buf", (size_t)65536 * 32 + 4096);
    if (!buf) return ctx->fail(NZCB_E_NOMEM, "microbench: out of memory");
    for (int rep = 0; rep < 2; rep++) {
        NZ_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
        NZ_LAUNCH(ctx, k_level_roundtrip, 1, 256, 0, buf, iters, kind, buf + 65536 * 8);
        NZ_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
        NZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    cudaEventElapsedTime(&ctx->last_ms, ctx->ev0, ctx->ev1);
    *ns_per_level = ctx->last_ms * 1e6 / iters;
    return 0;
}
