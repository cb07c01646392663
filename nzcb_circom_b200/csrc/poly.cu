// Segmented-scan restatements of snarkjs' sequential polynomial loops (see poly.cuh).
#include "poly.cuh"

namespace nzcb {

// out[s] = sum_{k<SEG} in[s*SEG+k] * x^k
__global__ void __launch_bounds__(128) k_horner_up(const Fr* __restrict__ in, size_t n_in, Fr x, Fr* __restrict__ out,
                                                   size_t n_out) {
    const size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_out) return;
    const size_t lo = s * NZ_SEG;
    size_t hi = lo + NZ_SEG;
    if (hi > n_in) hi = n_in;
    Fr acc = Fr::zero();
    for (size_t i = hi; i > lo; i--) acc = acc * x + in[i - 1];
    out[s] = acc;
}

// H[i] = in[i] + x * H[i+1] inside segment s, seeded with H_above[s+1] (= H[(s+1)*SEG]).
// shift == 0: out[i] = H[i].   shift == 1: out[i-1] = H[i] (quotient), out[n_in-1] = 0.
__global__ void __launch_bounds__(128) k_horner_down(const Fr* __restrict__ in, size_t n_in, Fr x,
                                                     const Fr* __restrict__ h_above, size_t n_above,
                                                     Fr* __restrict__ out, int shift) {
    const size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t lo = s * NZ_SEG;
    if (lo >= n_in) return;
    size_t hi = lo + NZ_SEG;
    if (hi > n_in) hi = n_in;
    Fr run = (s + 1 < n_above) ? h_above[s + 1] : Fr::zero();
    if (shift && hi == n_in) out[n_in - 1] = Fr::zero();
    for (size_t i = hi; i > lo; i--) {
        run = in[i - 1] + x * run;
        if (shift) {
            if (i - 1 >= 1) out[i - 2] = run;
        } else {
            out[i - 1] = run;
        }
    }
}

int poly_horner(nzcb_ctx* ctx, const Fr* d_p, size_t n, const Fr& x, Fr* d_value, Fr* d_quot) {
    if (n == 0) return ctx->fail(NZCB_E_INVALID, "poly_horner: empty polynomial");
    // level sizes
    std::vector<size_t> len;
    len.push_back(n);
    while (len.back() > 1) len.push_back((len.back() + NZ_SEG - 1) / NZ_SEG);
    const int top = (int)len.size() - 1;  // level `top` has length 1
    size_t total = 0;
    for (int l = 1; l <= top; l++) total += len[l];
    Fr* up = (Fr*)ctx->scratch_get("horner_up", (total + 1) * sizeof(Fr));
    Fr* dn = (Fr*)ctx->scratch_get("horner_dn", (total + 1) * sizeof(Fr));
    if (!up || !dn) return ctx->fail(NZCB_E_NOMEM, "poly_horner: out of device memory");
    std::vector<Fr*> a(top + 1), h(top + 1);
    std::vector<Fr> xs(top + 1);
    a[0] = const_cast<Fr*>(d_p);
    h[0] = nullptr;
    size_t off = 0;
    xs[0] = x;
    for (int l = 1; l <= top; l++) {
        a[l] = up + off;
        h[l] = dn + off;
        off += len[l];
        Fr t = xs[l - 1];
        for (int k = 0; k < 6; k++) t = t.sqr();  // x^(SEG) , SEG = 64
        xs[l] = t;
    }
    if (top == 0) {
        // single coefficient: value = p[0], quotient = 0
        NZ_CUDA(ctx, cudaMemcpyAsync(d_value, d_p, sizeof(Fr), cudaMemcpyDeviceToDevice, ctx->stream));
        if (d_quot) NZ_CUDA(ctx, cudaMemsetAsync(d_quot, 0, sizeof(Fr), ctx->stream));
        return 0;
    }
    for (int l = 0; l < top; l++)
        NZ_LAUNCH(ctx, k_horner_up, div_up(len[l + 1], 128), 128, 0, a[l], len[l], xs[l], a[l + 1], len[l + 1]);
    NZ_CUDA(ctx, cudaMemcpyAsync(d_value, a[top], sizeof(Fr), cudaMemcpyDeviceToDevice, ctx->stream));
    if (!d_quot) return 0;
    // top-down: H^top = a^top
    NZ_CUDA(ctx, cudaMemcpyAsync(h[top], a[top], sizeof(Fr), cudaMemcpyDeviceToDevice, ctx->stream));
    for (int l = top - 1; l >= 0; l--) {
        Fr* out = (l == 0) ? d_quot : h[l];
        NZ_LAUNCH(ctx, k_horner_down, div_up(len[l + 1], 128), 128, 0, a[l], len[l], xs[l], h[l + 1], len[l + 1], out,
                  l == 0 ? 1 : 0);
    }
    return 0;
}

// Several evaluations at once (round 4 evaluates seven polynomials): the up-sweeps of all of them share one launch per
// level (blockIdx.y = polynomial) instead of running one latency-bound chain of launches after the other.
struct HornerMulti {
    const Fr* in[NZ_HORNER_MAX];
    Fr* out[NZ_HORNER_MAX];
    size_t n_in[NZ_HORNER_MAX], n_out[NZ_HORNER_MAX];
    Fr x[NZ_HORNER_MAX];
};
__global__ void __launch_bounds__(128) k_horner_up_multi(HornerMulti a) {
    const int k = blockIdx.y;
    const size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= a.n_out[k]) return;
    const size_t lo = s * NZ_SEG;
    size_t hi = lo + NZ_SEG;
    if (hi > a.n_in[k]) hi = a.n_in[k];
    const Fr* in = a.in[k];
    const Fr x = a.x[k];
    Fr acc = Fr::zero();
    for (size_t i = hi; i > lo; i--) acc = acc * x + in[i - 1];
    a.out[k][s] = acc;
}

int poly_horner_multi(nzcb_ctx* ctx, int K, const Fr* const* d_p, const size_t* n, const Fr* x, Fr* const* d_value) {
    if (K < 1 || K > NZ_HORNER_MAX) return ctx->fail(NZCB_E_INVALID, "poly_horner_multi: 1..%d polynomials", NZ_HORNER_MAX);
    // level sizes per polynomial; all scratch in one arena
    std::vector<std::vector<size_t>> len(K);
    size_t total = 0;
    int top_max = 0;
    for (int k = 0; k < K; k++) {
        if (n[k] == 0) return ctx->fail(NZCB_E_INVALID, "poly_horner_multi: empty polynomial");
        len[k].push_back(n[k]);
        while (len[k].back() > 1) len[k].push_back((len[k].back() + NZ_SEG - 1) / NZ_SEG);
        for (size_t l = 1; l < len[k].size(); l++) total += len[k][l];
        top_max = std::max(top_max, (int)len[k].size() - 1);
    }
    Fr* up = (Fr*)ctx->scratch_get("horner_multi_up", (total + 1) * sizeof(Fr));
    if (!up) return ctx->fail(NZCB_E_NOMEM, "poly_horner_multi: out of device memory");
    std::vector<const Fr*> cur(K);
    std::vector<Fr> xs(x, x + K);
    std::vector<int> level(K, 0);
    for (int k = 0; k < K; k++) cur[k] = d_p[k];
    size_t off = 0;
    for (int l = 0; l < top_max; l++) {
        HornerMulti a;
        memset(&a, 0, sizeof(a));
        size_t widest = 0;
        for (int k = 0; k < K; k++) {
            const int top = (int)len[k].size() - 1;
            if (l >= top) {  // this polynomial is done: nothing to do at this level
                a.n_out[k] = 0;
                continue;
            }
            a.in[k] = cur[k];
            a.n_in[k] = len[k][l];
            a.n_out[k] = len[k][l + 1];
            a.out[k] = up + off;
            a.x[k] = xs[k];
            off += len[k][l + 1];
            widest = std::max(widest, a.n_out[k]);
        }
        NZ_LAUNCH(ctx, k_horner_up_multi, dim3(div_up(widest, 128), (unsigned)K), 128, 0, a);
        for (int k = 0; k < K; k++) {
            if (a.n_out[k] == 0) continue;
            cur[k] = a.out[k];
            Fr t = xs[k];
            for (int q = 0; q < 6; q++) t = t.sqr();  // x^(SEG), SEG = 64
            xs[k] = t;
        }
    }
    for (int k = 0; k < K; k++)  // single-coefficient polynomials: value = p[0]
        NZ_CUDA(ctx, cudaMemcpyAsync(d_value[k], cur[k], sizeof(Fr), cudaMemcpyDeviceToDevice, ctx->stream));
    return 0;
}

__global__ void __launch_bounds__(128) k_prod_up(const Fr* __restrict__ in, size_t n_in, Fr* __restrict__ out, size_t n_out) {
    const size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_out) return;
    const size_t lo = s * NZ_SEG;
    size_t hi = lo + NZ_SEG;
    if (hi > n_in) hi = n_in;
    Fr acc = in[lo];
    for (size_t i = lo + 1; i < hi; i++) acc = acc * in[i];
    out[s] = acc;
}

// out[i] = e_above[s] * prod_{lo<=j<i} in[j]   (alias-safe for out == in)
__global__ void __launch_bounds__(128) k_prod_down(const Fr* in, size_t n_in, const Fr* __restrict__ e_above, Fr* out) {
    const size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t lo = s * NZ_SEG;
    if (lo >= n_in) return;
    size_t hi = lo + NZ_SEG;
    if (hi > n_in) hi = n_in;
    Fr run = e_above[s];
    for (size_t i = lo; i < hi; i++) {
        const Fr v = in[i];
        out[i] = run;
        run = run * v;
    }
}

__global__ void k_set_one(Fr* p) { *p = Fr::one(); }

int prefix_product(nzcb_ctx* ctx, const Fr* d_in, size_t n, Fr* d_out, Fr* d_total) {
    if (n == 0) return ctx->fail(NZCB_E_INVALID, "prefix_product: empty input");
    std::vector<size_t> len;
    len.push_back(n);
    while (len.back() > 1) len.push_back((len.back() + NZ_SEG - 1) / NZ_SEG);
    const int top = (int)len.size() - 1;
    size_t total = 0;
    for (int l = 1; l <= top; l++) total += len[l];
    Fr* up = (Fr*)ctx->scratch_get("prod_up", (total + 1) * sizeof(Fr));
    Fr* dn = (Fr*)ctx->scratch_get("prod_dn", (total + 1) * sizeof(Fr));
    if (!up || !dn) return ctx->fail(NZCB_E_NOMEM, "prefix_product: out of device memory");
    std::vector<const Fr*> a(top + 1);
    std::vector<Fr*> e(top + 1);
    a[0] = d_in;
    e[0] = d_out;
    size_t off = 0;
    for (int l = 1; l <= top; l++) {
        a[l] = up + off;
        e[l] = dn + off;
        off += len[l];
    }
    if (top == 0) {
        NZ_CUDA(ctx, cudaMemcpyAsync(d_total, d_in, sizeof(Fr), cudaMemcpyDeviceToDevice, ctx->stream));
        NZ_LAUNCH(ctx, k_set_one, 1, 1, 0, d_out);
        return 0;
    }
    for (int l = 0; l < top; l++)
        NZ_LAUNCH(ctx, k_prod_up, div_up(len[l + 1], 128), 128, 0, a[l], len[l], const_cast<Fr*>(a[l + 1]), len[l + 1]);
    NZ_CUDA(ctx, cudaMemcpyAsync(d_total, a[top], sizeof(Fr), cudaMemcpyDeviceToDevice, ctx->stream));
    NZ_LAUNCH(ctx, k_set_one, 1, 1, 0, e[top]);
    for (int l = top - 1; l >= 0; l--)
        NZ_LAUNCH(ctx, k_prod_down, div_up(len[l + 1], 128), 128, 0, a[l], len[l], e[l + 1], e[l]);
    return 0;
}

constexpr int NZ_INV_CHUNK = 32;
__global__ void __launch_bounds__(128) k_batch_inverse(Fr* __restrict__ a, Fr* __restrict__ tmp, size_t n) {
    const size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t lo = s * NZ_INV_CHUNK;
    if (lo >= n) return;
    size_t hi = lo + NZ_INV_CHUNK;
    if (hi > n) hi = n;
    Fr acc = a[lo];
    tmp[lo] = acc;
    for (size_t i = lo + 1; i < hi; i++) {
        acc = acc * a[i];
        tmp[i] = acc;
    }
    Fr inv = acc.inv();
    for (size_t i = hi - 1; i > lo; i--) {
        const Fr v = a[i];
        a[i] = inv * tmp[i - 1];
        inv = inv * v;
    }
    a[lo] = inv;
}

int batch_inverse(nzcb_ctx* ctx, Fr* d_a, size_t n) {
    if (n == 0) return 0;
    Fr* tmp = (Fr*)ctx->scratch_get("batch_inv_tmp", n * sizeof(Fr));
    if (!tmp) return ctx->fail(NZCB_E_NOMEM, "batch_inverse: out of device memory");
    NZ_LAUNCH(ctx, k_batch_inverse, div_up(div_up(n, NZ_INV_CHUNK), 128), 128, 0, d_a, tmp, n);
    return 0;
}

}  // namespace nzcb
