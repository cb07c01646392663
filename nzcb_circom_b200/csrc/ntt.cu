// Fr NTT / iNTT over the 2^k domain -- replaces ffjavascript Fr.fft / Fr.ifft
// (un-vendored, /root/reference/yarn.lock:3905; semantics SURVEY.md A.1:
// natural order in, natural order out, ifft carries the 1/N factor).
//
// Decimation-in-frequency.  One launch fuses up to three radix-2 stages: a
// thread owns 2^R elements that are N>>(s+R) apart, so a warp's loads of each
// "row" are 32 consecutive 32-byte elements (1 KiB, fully coalesced) for all
// but the last two passes, where every thread reads whole 32-byte sectors of
// its own contiguous run.  Twiddles come from a per-size table w^t, t < N/2,
// read with unit stride in the early stages (coalesced LDG.128 pairs) and
// served from L1/L2 in the late ones.  A final pass undoes the bit reversal
// (and folds in 1/N for the inverse).
//
// Roofline: (N/2) log2 N modmul = 132 N log2 N IMAD32; HBM traffic
// is 64 N bytes per pass, ceil(log2 N / 3) + 1 passes -- integer-pipe bound.
#include "common.cuh"
#include "poly.cuh"

namespace nzcb {

__global__ void k_twiddle_fill(Fr* __restrict__ W, Fr w, size_t half) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= half) return;
    W[t] = w.pow_u64(t);
}

template <int R>
__global__ void __launch_bounds__(256) k_ntt_pass(Fr* __restrict__ a, const Fr* __restrict__ W, uint32_t log_n,
                                                   uint32_t s) {
    constexpr int M = 1 << R;
    const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t log_q = log_n - s - R;
    if (g >= ((size_t)1 << (log_n - R))) return;
    const size_t q = (size_t)1 << log_q;
    const size_t j = g & (q - 1);
    const size_t p = ((g >> log_q) << (log_q + R)) + j;

    Fr x[M];
#pragma unroll
    for (int m = 0; m < M; m++) x[m] = a[p + (size_t)m * q];

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
#pragma unroll
    for (int m = 0; m < M; m++) a[p + (size_t)m * q] = x[m];
}

// in-place bit-reversal permutation; scale != nullptr multiplies every element by *scale
__global__ void k_bitrev(Fr* __restrict__ a, uint32_t log_n, Fr scale, int do_scale) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ((size_t)1 << log_n)) return;
    const size_t r = log_n ? (size_t)(__brev((uint32_t)i) >> (32 - log_n)) : 0;
    if (i < r) {
        Fr u = a[i], v = a[r];
        if (do_scale) {
            u = u * scale;
            v = v * scale;
        }
        a[i] = v;
        a[r] = u;
    } else if (i == r && do_scale) {
        a[i] = a[i] * scale;
    }
}

// the same permutation with a per-element factor: a'[i] = a[rev(i)] * tab[i]   (coset iNTT: tab[i] = g^-i / N)
__global__ void k_bitrev_tab(Fr* __restrict__ a, uint32_t log_n, const Fr* __restrict__ tab) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ((size_t)1 << log_n)) return;
    const size_t r = log_n ? (size_t)(__brev((uint32_t)i) >> (32 - log_n)) : 0;
    if (i < r) {
        const Fr u = a[i], v = a[r];
        a[i] = v * tab[i];
        a[r] = u * tab[r];
    } else if (i == r) {
        a[i] = a[i] * tab[i];
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
    uint32_t s = 0;
    const uint32_t rem = log_n % 3;
    if (rem == 1) {
        NZ_LAUNCH(ctx, k_ntt_pass<1>, div_up((size_t)1 << (log_n - 1), 256), 256, 0, d, W, log_n, s);
        s += 1;
    } else if (rem == 2) {
        NZ_LAUNCH(ctx, k_ntt_pass<2>, div_up((size_t)1 << (log_n - 2), 256), 256, 0, d, W, log_n, s);
        s += 2;
    }
    for (; s < log_n; s += 3) {
        NZ_LAUNCH(ctx, k_ntt_pass<3>, div_up((size_t)1 << (log_n - 3), 256), 256, 0, d, W, log_n, s);
    }
    Fr scale = Fr::one();
    if (inverse) scale = Fr::from_u64((uint64_t)1 << log_n).inv();
    if (tab) NZ_LAUNCH(ctx, k_bitrev_tab, div_up((size_t)1 << log_n, 256), 256, 0, d, log_n, tab);
    else NZ_LAUNCH(ctx, k_bitrev, div_up((size_t)1 << log_n, 256), 256, 0, d, log_n, scale, inverse ? 1 : 0);
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
