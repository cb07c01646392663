// Lagrange-basis SRS: [L_i(tau)]G1 for the size-n domain from the monomial points [tau^j]G1 of zkey section 14
// -- the group inverse DFT that `snarkjs powersoftau prepare phase2` performs (ptau sections 12-15,
// SURVEY.md A.4); a PLONK zkey does not carry those points, so they are derived once per key at load.
//
// Why: round 1 commits to a, b, c.  In the monomial basis their coefficients are full 254-bit scalars; in the
// Lagrange basis the scalars are the wire values themselves -- 0, +-1, bytes, 32-bit words for ~97 % of the rows
// of nzcp_live -- so  [a(tau)] = sum_i A_i [L_i(tau)] + b1 ([tau^(n+1)] - [tau]) + b2 ([tau^n] - [1])  is the same
// group element for ~5 % of the bucket additions.
//
// L_i = (1/n) sum_j w^(-ij) P_j : radix-2 decimation-in-frequency butterflies on XYZZ points,
// (u, v) -> (u + v, w^-e (u - v)), one launch per stage, then a bit-reversal pass that multiplies by 1/n and
// converts to affine.  The scalar multiplications use signed 4-bit windows.
#include "common.cuh"
#include "poly.cuh"

namespace nzcb {

// k * P for a canonical 256-bit scalar k < r: signed radix-16 digits, table 1P..8P in registers / local memory
__device__ __noinline__ G1XYZZ g1_mul_scalar(const G1XYZZ& P, const Fr& k_canonical) {
    if (P.is_inf() || k_canonical.is_zero()) return G1XYZZ::inf();
    G1XYZZ tab[8];
    tab[0] = P;
    tab[1] = P.dbl();
#pragma unroll 1
    for (int i = 2; i < 8; i++) {
        tab[i] = tab[i - 1];
        tab[i].add(P);
    }
    // recode into 64 signed digits in [-8, 8], least significant first; r < 2^254 so the top digit absorbs the carry
    int8_t dig[65];
    uint32_t carry = 0;
#pragma unroll 1
    for (int i = 0; i < 64; i++) {
        uint32_t d = ((k_canonical.v[i >> 3] >> ((i & 7) * 4)) & 15u) + carry;
        if (d > 8) {
            dig[i] = (int8_t)((int)d - 16);
            carry = 1;
        } else {
            dig[i] = (int8_t)d;
            carry = 0;
        }
    }
    dig[64] = (int8_t)carry;
    G1XYZZ acc = G1XYZZ::inf();
#pragma unroll 1
    for (int i = 64; i >= 0; i--) {
        if (i != 64) {
            acc = acc.dbl();
            acc = acc.dbl();
            acc = acc.dbl();
            acc = acc.dbl();
        }
        const int d = dig[i];
        if (d > 0) {
            acc.add(tab[d - 1]);
        } else if (d < 0) {
            acc.add(tab[-d - 1].neg());
        }
    }
    return acc;
}

__global__ void __launch_bounds__(128) k_g1_from_affine(const G1Affine* __restrict__ in, G1XYZZ* __restrict__ out, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = G1XYZZ::from_affine(in[i]);
}

// one DIF stage s: half = n >> (s + 1); butterfly t = (block, j): twiddle w_inv^(j << s), table Winv[t], t < n/2
__global__ void __launch_bounds__(128) k_g1_fft_stage(G1XYZZ* __restrict__ a, const Fr* __restrict__ Winv, uint32_t log_n,
                                                      uint32_t s) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (1u << (log_n - 1))) return;
    const uint32_t log_half = log_n - s - 1;
    const uint32_t half = 1u << log_half;
    const uint32_t j = t & (half - 1);
    const uint32_t i0 = ((t >> log_half) << (log_half + 1)) + j;
    const uint32_t i1 = i0 + half;
    const G1XYZZ u = a[i0], v = a[i1];
    G1XYZZ sum = u;
    sum.add(v);
    G1XYZZ diff = u;
    diff.add(v.neg());
    a[i0] = sum;
    const uint32_t e = j << s;
    a[i1] = e == 0 ? diff : g1_mul_scalar(diff, Winv[e].from_mont());
}

// out[i] = affine( (1/n) * a[bitrev(i)] )
__global__ void __launch_bounds__(128) k_g1_fft_finish(const G1XYZZ* __restrict__ a, uint32_t log_n, Fr n_inv_canonical,
                                                       G1Affine* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (1u << log_n)) return;
    const uint32_t r = log_n ? (__brev(i) >> (32 - log_n)) : 0;
    out[i] = g1_mul_scalar(a[r], n_inv_canonical).to_affine();
}

int g1_lagrange_basis(nzcb_ctx* ctx, const G1Affine* d_srs, uint32_t log_n, G1Affine* d_out) {
    const size_t n = (size_t)1 << log_n;
    G1XYZZ* buf = (G1XYZZ*)ctx->scratch_get("g1fft_buf", n * sizeof(G1XYZZ));
    if (!buf) return ctx->fail(NZCB_E_NOMEM, "lagrange basis: cannot allocate %zu bytes", n * sizeof(G1XYZZ));
    NZ_LAUNCH(ctx, k_g1_from_affine, div_up(n, 128), 128, 0, d_srs, buf, (uint32_t)n);
    if (log_n) {
        const Fr* Winv = nullptr;
        NZ_TRY(get_twiddles_pub(ctx, log_n, true, &Winv));
        for (uint32_t s = 0; s < log_n; s++)
            NZ_LAUNCH(ctx, k_g1_fft_stage, div_up(n / 2, 128), 128, 0, buf, Winv, log_n, s);
    }
    const Fr n_inv = Fr::from_u64(n).inv().from_mont();
    NZ_LAUNCH(ctx, k_g1_fft_finish, div_up(n, 128), 128, 0, buf, log_n, n_inv, d_out);
    return 0;
}

}  // namespace nzcb

using namespace nzcb;

// [L_i(tau)]G1, i < 2^log_n, from [tau^j]G1, j < 2^log_n (both n x 64 B affine LEM, host buffers)
extern "C" int32_t nzcb_g1_lagrange_basis(nzcb_ctx* ctx, const uint8_t* srs_affine_lem, uint32_t log_n, uint8_t* out_affine_lem) {
    if (!ctx || !srs_affine_lem || !out_affine_lem || log_n > 26) return NZCB_E_INVALID;
    NZ_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t n = (size_t)1 << log_n;
    G1Affine* d_in = (G1Affine*)ctx->scratch_get("g1fft_in", n * sizeof(G1Affine));
    G1Affine* d_out = (G1Affine*)ctx->scratch_get("g1fft_out", n * sizeof(G1Affine));
    if (!d_in || !d_out) return ctx->fail(NZCB_E_NOMEM, "lagrange basis: out of device memory");
    NZ_CUDA(ctx, cudaMemcpyAsync(d_in, srs_affine_lem, n * 64, cudaMemcpyHostToDevice, ctx->stream));
    NZ_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
    NZ_TRY(g1_lagrange_basis(ctx, d_in, log_n, d_out));
    NZ_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
    NZ_CUDA(ctx, cudaMemcpyAsync(out_affine_lem, d_out, n * 64, cudaMemcpyDeviceToHost, ctx->stream));
    NZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    cudaEventElapsedTime(&ctx->last_ms, ctx->ev0, ctx->ev1);
    return 0;
}
