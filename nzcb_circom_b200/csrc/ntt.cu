// Fr NTT / iNTT over the 2^k domain -- replaces ffjavascript Fr.fft / Fr.ifft
// (un-vendored, /root/reference/yarn.lock:3905; semantics SURVEY.md A.1:
// natural order in, natural order out, ifft carries the 1/N factor).
//
// Decimation-in-frequency.  One launch fuses up to three radix-2 stages: a
// thread owns 2^R elements that are N>>(s+R) apart, so a warp's loads of each
// "row" are 32 consecutive 32-byte elements (1 KiB, fully coalesced) for all
// but the last two passes, where every thread reads whole 32-byte sectors of
// its own contiguous run.  Twiddles come from a per-size table w^t, t < N/2,
// read with unit stride in the early stages and served from L1/L2 in the late
// ones.  The first launch writes a ping-pong buffer, the last one scatters to
// bit-reversed (= natural) positions in the caller's array and applies 1/N or
// the caller's per-element factors (coset transforms) -- no separate
// permutation pass.
//
// Roofline: (N/2) log2 N modmul = 132 N log2 N IMAD32; HBM traffic
// is 64 N bytes per launch, ceil(log2 N / 3) launches -- integer-pipe bound.
// A shared-memory variant (6-8 stages per launch, warp-private transposes) was
// measured slower at 2^23 (3.6 vs 2.9 ms: its 36 inlined multiplies thrash the
// instruction cache, ncu stall_no_instruction 8.6) and was dropped.
#include "common.cuh"
#include "poly.cuh"
#include <stdlib.h>

namespace nzcb {

__global__ void k_twiddle_fill(Fr* __restrict__ W, Fr w, size_t half) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= half) return;
    W[t] = w.pow_u64(t);
}

// One launch = R radix-2 stages [s, s+R) in registers.  `last`: the results go to their bit-reversed positions
// (natural-order output) multiplied by tab[position] or by `scale` -- so a transform is ceil(log2 N / 3) launches,
// first one out of place into a ping-pong buffer, last one back into the caller's array.
template <int R, int TH, int MINB>
__global__ void __launch_bounds__(TH, MINB) k_ntt_pass(const Fr* in, Fr* out, const Fr* __restrict__ W, uint32_t log_n,
                                                       uint32_t s, int last, Fr scale, int do_scale,
                                                       const Fr* __restrict__ tab) {
    constexpr int M = 1 << R;
    const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t log_q = log_n - s - R;
    if (g >= ((size_t)1 << (log_n - R))) return;
    const size_t q = (size_t)1 << log_q;
    const size_t j = g & (q - 1);
    const size_t p = ((g >> log_q) << (log_q + R)) + j;

    Fr x[M];
#pragma unroll
    for (int m = 0; m < M; m++) x[m] = in[p + (size_t)m * q];

#pragma unroll
    for (int t = 0; t < R; t++) {
        const int dm = 1 << (R - 1 - t);
#pragma unroll
        for (int m = 0; m < M; m++) {
            if ((m / dm) & 1) continue;  // m is the upper leg of a butterfly
            const size_t e = (j + (size_t)(m % dm) * q) << (s + t);
            const Fr w = W[e];
            const Fr u = x[m];
            const Fr v = x[m + dm];
            x[m] = u + v;
            x[m + dm] = (u - v) * w;
        }
    }
    if (!last) {
#pragma unroll
        for (int m = 0; m < M; m++) out[p + (size_t)m * q] = x[m];
    } else {
#pragma unroll
        for (int m = 0; m < M; m++) {
            const size_t idx = p + (size_t)m * q;
            const size_t r = (size_t)(__brev((uint32_t)idx) >> (32 - log_n));
            Fr o = x[m];
            if (tab) o = o * tab[r];
            else if (do_scale) o = o * scale;
            out[r] = o;
        }
    }
}

// 2^log_n-th primitive root w[log_n]: w[28] = 5^((r-1)/2^28), w[i] = w[i+1]^2
Fr fr_root_host(uint32_t log_n) {
    // (r - 1) >> 28
    uint32_t e[8];
    for (int i = 0; i < 8; i++) e[i] = FrParams::mod(i);
    e[0] -= 1;
    uint32_t sh[8];
    for (int i = 0; i < 8; i++) {
        uint64_t lo = e[i] >> 28;
        uint64_t hi = (i < 7) ? ((uint64_t)e[i + 1] << 4) : 0;
        sh[i] = (uint32_t)(lo | hi);
    }
    Fr w = Fr::from_u64(5).pow_limbs(sh);
    for (uint32_t i = 28; i > log_n; i--) w = w.sqr();
    return w;
}

// tables live in the root ctx and are shared by its lanes (built once, complete before anyone reads them)
int get_twiddles_pub(nzcb_ctx* ctx, uint32_t log_n, bool inverse, const Fr** out) {
    nzcb_ctx* root = ctx->root();
    std::lock_guard<std::mutex> g(root->mu);
    const uint32_t key = log_n * 2 + (inverse ? 1 : 0);
    auto it = root->twiddles.find(key);
    if (it != root->twiddles.end()) {
        *out = it->second;
        return 0;
    }
    const size_t half = log_n ? ((size_t)1 << (log_n - 1)) : 1;
    Fr* W = nullptr;
    NZ_CUDA(ctx, cudaMalloc(&W, half * sizeof(Fr)));
    Fr w = fr_root_host(log_n);
    if (inverse) w = w.inv();
    NZ_LAUNCH(ctx, k_twiddle_fill, div_up(half, 256), 256, 0, W, w, half);
    NZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    root->twiddles[key] = W;
    *out = W;
    return 0;
}

int ntt_dev(nzcb_ctx* ctx, Fr* d, uint32_t log_n, bool inverse) { return ntt_dev_tab(ctx, d, log_n, inverse, nullptr); }

// tab != nullptr: the final pass multiplies output element i by tab[i] INSTEAD of the uniform 1/N of the inverse
int ntt_dev_tab(nzcb_ctx* ctx, Fr* d, uint32_t log_n, bool inverse, const Fr* tab) {
    if (log_n > 28) return ctx->fail(NZCB_E_INVALID, "ntt: log_n %u exceeds the 2-adicity of Fr (28)", log_n);
    if (log_n == 0) return 0;
    const Fr* W = nullptr;
    NZ_TRY(get_twiddles_pub(ctx, log_n, inverse, &W));
    Fr scale = Fr::one();
    if (inverse) scale = Fr::from_u64((uint64_t)1 << log_n).inv();
    Fr* tmp = (Fr*)ctx->scratch_get("ntt_tmp", ((size_t)1 << log_n) * sizeof(Fr));
    if (!tmp) return ctx->fail(NZCB_E_NOMEM, "ntt: cannot allocate the ping-pong buffer");
    static const int variant = [] {
        const char* e = getenv("NZCB_NTT_VAR");
        return e ? atoi(e) : 2;
    }();
    const uint32_t rem = log_n % 3;
    const uint32_t n_pass = log_n / 3 + (rem ? 1 : 0);
    uint32_t s = 0;
    for (uint32_t pi = 0; pi < n_pass; pi++) {
        const uint32_t R = (pi == 0 && rem) ? rem : 3;
        const int last = pi + 1 == n_pass;
        const Fr* src = pi == 0 ? d : tmp;
        Fr* dst = (last && n_pass > 1) ? d : tmp;
        const size_t n_thr = (size_t)1 << (log_n - R);
        const int dsc = inverse ? 1 : 0;
        if (R == 1) NZ_LAUNCH(ctx, (k_ntt_pass<1, 256, 1>), div_up(n_thr, 256), 256, 0, src, dst, W, log_n, s, last, scale, dsc, tab);
        else if (R == 2) NZ_LAUNCH(ctx, (k_ntt_pass<2, 256, 1>), div_up(n_thr, 256), 256, 0, src, dst, W, log_n, s, last, scale, dsc, tab);
        else if (variant == 0) NZ_LAUNCH(ctx, (k_ntt_pass<3, 256, 1>), div_up(n_thr, 256), 256, 0, src, dst, W, log_n, s, last, scale, dsc, tab);
        else if (variant == 2) NZ_LAUNCH(ctx, (k_ntt_pass<3, 256, 2>), div_up(n_thr, 256), 256, 0, src, dst, W, log_n, s, last, scale, dsc, tab);
        else if (variant == 3) NZ_LAUNCH(ctx, (k_ntt_pass<3, 128, 3>), div_up(n_thr, 128), 128, 0, src, dst, W, log_n, s, last, scale, dsc, tab);
        else NZ_LAUNCH(ctx, (k_ntt_pass<3, 128, 4>), div_up(n_thr, 128), 128, 0, src, dst, W, log_n, s, last, scale, dsc, tab);
        s += R;
    }
    if (n_pass == 1)  // a single launch cannot permute in place: it wrote tmp
        NZ_CUDA(ctx, cudaMemcpyAsync(d, tmp, ((size_t)1 << log_n) * sizeof(Fr), cudaMemcpyDeviceToDevice, ctx->stream));
    return 0;
}

}  // namespace nzcb

using namespace nzcb;

extern "C" int32_t nzcb_ntt_fr_dev(nzcb_ctx* ctx, void* d_data, uint32_t log_n, int32_t inverse) {
    if (!ctx || !d_data) return NZCB_E_INVALID;
    NZ_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
    NZ_TRY(ntt_dev(ctx, (Fr*)d_data, log_n, inverse != 0));
    NZ_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
    NZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    cudaEventElapsedTime(&ctx->last_ms, ctx->ev0, ctx->ev1);
    return 0;
}

extern "C" int32_t nzcb_ntt_fr(nzcb_ctx* ctx, uint8_t* data, uint32_t log_n, int32_t inverse) {
    if (!ctx || !data) return NZCB_E_INVALID;
    if (log_n > 28) return ctx->fail(NZCB_E_INVALID, "ntt: log_n %u exceeds the 2-adicity of Fr (28)", log_n);
    const size_t bytes = ((size_t)1 << log_n) * sizeof(Fr);
    Fr* d = (Fr*)ctx->scratch_get("ntt_io", bytes);
    if (!d) return ctx->fail(NZCB_E_NOMEM, "ntt: cannot allocate %zu device bytes", bytes);
    NZ_CUDA(ctx, cudaMemcpyAsync(d, data, bytes, cudaMemcpyHostToDevice, ctx->stream));
    NZ_TRY(ntt_dev(ctx, d, log_n, inverse != 0));
    NZ_CUDA(ctx, cudaMemcpyAsync(data, d, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    NZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}
