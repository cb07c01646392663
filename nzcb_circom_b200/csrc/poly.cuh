// Polynomial helpers shared by the prover and the setup: the sequential loops of
// snarkjs (evalPol, divPol1, the grand-product loop, batchInverse -- SURVEY.md
// A.2 rounds 2, 4, 5) restated as segmented scans.  Field arithmetic is exact,
// so any algebraically equal schedule gives bit-identical results.
#pragma once
#include "common.cuh"

namespace nzcb {

constexpr int NZ_SEG = 64;  // elements one thread walks sequentially in a scan level

// Horner over a polynomial, bottom-up in segments of NZ_SEG:
//   H(k) = sum_{j>=k} p[j] x^(j-k).   value = H(0) = p(x)   (snarkjs evalPol)
//   quotient q[i] = H(i+1), q[n-1] = 0                       (snarkjs divPol1: p / (X - x))
// d_value: device Fr receiving p(x);  d_quot: nullptr or n elements.
int poly_horner(nzcb_ctx* ctx, const Fr* d_p, size_t n, const Fr& x, Fr* d_value, Fr* d_quot);

// K <= NZ_HORNER_MAX evaluations p_k(x_k) with one launch per level for all of them (values only, no quotient)
constexpr int NZ_HORNER_MAX = 8;
int poly_horner_multi(nzcb_ctx* ctx, int K, const Fr* const* d_p, const size_t* n, const Fr* x, Fr* const* d_value);

// exclusive prefix product: out[i] = prod_{j<i} in[j]; *d_total = prod of all.  in == out allowed.
int prefix_product(nzcb_ctx* ctx, const Fr* d_in, size_t n, Fr* d_out, Fr* d_total);

// in-place batch inverse (Montgomery trick per 32-element chunk, one Fermat inverse each)
int batch_inverse(nzcb_ctx* ctx, Fr* d_a, size_t n);

// w^i for the size-2^log_n domain, from the NTT twiddle table (t < N/2) : w^(t + N/2) = -w^t
__device__ __forceinline__ Fr domain_pow(const Fr* __restrict__ W, uint32_t log_n, size_t i) {
    const size_t half = (size_t)1 << (log_n - 1);
    return i < half ? W[i] : W[i - half].neg();
}

int get_twiddles_pub(nzcb_ctx* ctx, uint32_t log_n, bool inverse, const Fr** out);
Fr fr_root_host(uint32_t log_n);

}  // namespace nzcb
