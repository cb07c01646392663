// PLONK prover -- replaces snarkjs 0.4.12 plonk.prove(zkey, wtns) (un-vendored,
// /root/reference/yarn.lock:7279; the reference only names it in its Makefile
// recipe, /root/reference/Makefile:54-62).  Data flow and transcript follow
// SURVEY.md A.2 step by step so that, with the nine blinders injected, the proof
// is the same 800 bytes snarkjs would print; the JS loops become kernels:
//   witness -> Montgomery, additions level by level, A/B/C gather       (round 1)
//   grand product: fused num/den, batch inverse, prefix-product scan     (round 2)
//   quotient T / Tz: one fused elementwise kernel over the 4n domain     (round 3)
//   evalPol / divPol1: segmented Horner scans (poly.cu)                  (rounds 4, 5)
// The zkey is parsed once and stays device resident in Montgomery form.
#include "common.cuh"
#include "keccak.h"
#include "poly.cuh"
#include <algorithm>
#include <atomic>
#include <thread>

using namespace nzcb;

struct nzcb_zkey {
    nzcb_ctx* ctx = nullptr;
    uint32_t n_vars = 0, n_public = 0, n = 0, power = 0, n_add = 0, n_cons = 0;
    Fr k1, k2;
    uint32_t* d_map[3] = {nullptr, nullptr, nullptr};
    // additions sorted by dependency level
    uint32_t *d_add_a = nullptr, *d_add_b = nullptr, *d_add_out = nullptr;
    Fr *d_add_ac = nullptr, *d_add_bc = nullptr;
    std::vector<uint32_t> level_off;  // additions of level l are [level_off[l], level_off[l+1])
    Fr* d_q[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};  // n coefs then 4n evals each
    Fr* d_sigma = nullptr;                                       // 3 x (n + 4n)
    Fr* d_lag = nullptr;                                         // max(nPublic,1) x (n + 4n)
    G1Table tab;                                                 // [tau^i]G1, i < n + 6, with window shifts (msm.cu)
    // round 1 commits in the Lagrange basis (g1fft.cu): [L_i(tau)]G1, i < n, then [1], [tau], [tau^n], [tau^(n+1)]
    // for the blinding terms -- n + 4 bases, the scalars are the wire values themselves
    G1Table tab_lag;
    // round 3 runs on the coset g * H_4n (Z_H never vanishes there): selectors, sigmas and the public-input
    // Lagrange polynomials evaluated on it once per key
    Fr g;                                                        // coset shift
    Fr zh_inv[4];                                                // 1 / (g^n w4^p - 1)
    Fr* d_cos[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // qm ql qr qo qc s1 s2 s3
    Fr* d_cos_lag = nullptr;                                     // max(nPublic,1) x 4n
    Fr* d_gpow = nullptr;                                        // g^k, k < n + 8
    Fr* d_ginv = nullptr;                                        // g^-k / 4n, k < 4n
};

namespace {

struct Section {
    const uint8_t* p;
    uint64_t size;
};

int parse_binfile(nzcb_ctx* ctx, const uint8_t* data, size_t len, const char* magic, std::map<uint32_t, Section>& out) {
    if (len < 12 || memcmp(data, magic, 4) != 0) return ctx->fail(NZCB_E_INVALID, "%s file: bad magic", magic);
    uint32_t nsec;
    memcpy(&nsec, data + 8, 4);
    size_t pos = 12;
    for (uint32_t i = 0; i < nsec; i++) {
        if (pos + 12 > len) return ctx->fail(NZCB_E_INVALID, "%s file: truncated section table", magic);
        uint32_t id;
        uint64_t size;
        memcpy(&id, data + pos, 4);
        memcpy(&size, data + pos + 4, 8);
        pos += 12;
        if (size > len - pos) return ctx->fail(NZCB_E_INVALID, "%s file: section %u overruns the file", magic, id);
        if (!out.count(id)) out[id] = Section{data + pos, size};
        pos += size;
    }
    return 0;
}

const uint8_t R_LE[32] = {0x01, 0x00, 0x00, 0xf0, 0x93, 0xf5, 0xe1, 0x43, 0x91, 0x70, 0xb9, 0x79, 0x48, 0xe8, 0x33, 0x28,
                          0x5d, 0x58, 0x81, 0x81, 0xb6, 0x45, 0x50, 0xb8, 0x29, 0xa0, 0x31, 0xe1, 0x72, 0x4e, 0x64, 0x30};
const uint8_t Q_LE[32] = {0x47, 0xfd, 0x7c, 0xd8, 0x16, 0x8c, 0x20, 0x3c, 0x8d, 0xca, 0x71, 0x68, 0x91, 0x6a, 0x81, 0x97,
                          0x5d, 0x58, 0x81, 0x81, 0xb6, 0x45, 0x50, 0xb8, 0x29, 0xa0, 0x31, 0xe1, 0x72, 0x4e, 0x64, 0x30};

template <class F>
void to_be_bytes(const F& mont, uint8_t out[32]) {
    F c = mont.from_mont();
    for (int i = 0; i < 8; i++)
        for (int k = 0; k < 4; k++) out[31 - (4 * i + k)] = (uint8_t)(c.v[i] >> (8 * k));
}
void g1_to_be(const G1Affine& p, uint8_t out[64]) {
    to_be_bytes(p.x, out);
    to_be_bytes(p.y, out + 32);
}
// snarkjs hashToFr: keccak256 digest as a big-endian integer, reduced mod r; returned in Montgomery form
Fr hash_to_fr(const std::vector<uint8_t>& msg) {
    uint8_t dg[32];
    keccak256(msg.data(), msg.size(), dg);
    Fr v;
    for (int i = 0; i < 8; i++) {
        uint32_t w = 0;
        for (int k = 0; k < 4; k++) w |= (uint32_t)dg[31 - (4 * i + k)] << (8 * k);
        v.v[i] = w;
    }
    // v < 2^256 < 6r: subtract r while v >= r
    for (int it = 0; it < 6; it++) {
        uint32_t t[8];
        uint32_t borrow = 0;
        for (int i = 0; i < 8; i++) {
            uint64_t d = (uint64_t)v.v[i] - FrParams::mod(i) - borrow;
            t[i] = (uint32_t)d;
            borrow = (uint32_t)(d >> 63);
        }
        if (borrow) break;
        for (int i = 0; i < 8; i++) v.v[i] = t[i];
    }
    return v.to_mont();
}
void append(std::vector<uint8_t>& v, const uint8_t* p, size_t n) { v.insert(v.end(), p, p + n); }

// ---------------------------------------------------------------- kernels
__global__ void k_wtns_to_mont(const Fr* __restrict__ w_le, Fr* __restrict__ W, size_t n_w) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_w) return;
    // "First element in plonk is not used ... We set it to zero" (A.2 step 0)
    W[i] = i == 0 ? Fr::zero() : fr_reduce_256(w_le[i]).to_mont();  // a value >= r in a .wtns is taken mod r
}

__global__ void k_additions(const uint32_t* __restrict__ ia, const uint32_t* __restrict__ ib,
                            const Fr* __restrict__ ac, const Fr* __restrict__ bc, const uint32_t* __restrict__ iout,
                            uint32_t lo, uint32_t hi, uint32_t n_vars, uint32_t n_w, Fr* __restrict__ W) {
    const uint32_t i = lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= hi) return;
    const uint32_t a = ia[i], b = ib[i];
    const Fr aw = a < n_vars ? W[a] : Fr::zero();
    const Fr bw = b < n_vars ? W[b] : Fr::zero();
    W[n_w + iout[i]] = ac[i] * aw + bc[i] * bw;
}

__global__ void k_gather(const uint32_t* __restrict__ map, const Fr* __restrict__ W, uint32_t n_vars, uint32_t n_cons,
                         uint32_t n, Fr* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fr v = Fr::zero();
    if (i < n_cons) {
        const uint32_t s = map[i];
        if (s < n_vars) v = W[s];
    }
    out[i] = v;
}

// pol (n + k coefficients) = coef + (pz[0] + pz[1] X + ...) * (X^n - 1)      (to4T, A.2 round 1)
struct Blind {
    Fr pz[3];
    int k;
};
__global__ void k_blind(Fr* __restrict__ pol, size_t n, Blind bl) {
    const int j = threadIdx.x;
    if (j >= bl.k) return;
    pol[n + j] = bl.pz[j];
    pol[j] = pol[j] - bl.pz[j];
}

struct R2Args {
    const Fr *A, *B, *C, *S14, *S24, *S34, *Wn;
    Fr beta, gamma, k1, k2;
    uint32_t n, power;
    Fr *num, *den;
};
__global__ void __launch_bounds__(128) k_round2_terms(R2Args a) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    const Fr w = domain_pow(a.Wn, a.power, i);
    const Fr bw = a.beta * w;
    const Fr av = a.A[i], bv = a.B[i], cv = a.C[i];
    const Fr n1 = av + bw + a.gamma;
    const Fr n2 = bv + a.k1 * bw + a.gamma;
    const Fr n3 = cv + a.k2 * bw + a.gamma;
    a.num[i] = n1 * (n2 * n3);
    const Fr d1 = av + a.S14[(size_t)4 * i] * a.beta + a.gamma;
    const Fr d2 = bv + a.S24[(size_t)4 * i] * a.beta + a.gamma;
    const Fr d3 = cv + a.S34[(size_t)4 * i] * a.beta + a.gamma;
    a.den[i] = d1 * (d2 * d3);
}

// the four blinding scalars of the Lagrange-basis commitment, after the n evaluations:
// pol = a + (pz0 + pz1 X)(X^n - 1)  ->  [1]: -pz0, [tau]: -pz1, [tau^n]: pz0, [tau^(n+1)]: pz1
__global__ void k_blind_scalars(Fr* __restrict__ evals, size_t n, Fr pz0, Fr pz1) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    evals[n] = pz0.neg();
    evals[n + 1] = pz1.neg();
    evals[n + 2] = pz0;
    evals[n + 3] = pz1;
}

__global__ void k_mul_inplace(Fr* __restrict__ a, const Fr* __restrict__ b, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    a[i] = a[i] * b[i];
}

// out[k] = scale * g^k
__global__ void k_pow_table(Fr* __restrict__ out, Fr g, Fr scale, size_t n) {
    const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    out[k] = g.pow_u64(k) * scale;
}
// coefficients of p(gX), zero-padded: out[k] = in[k] * g^k (k < m), 0 (m <= k < total)
__global__ void k_scale_pad(const Fr* __restrict__ in, size_t m, const Fr* __restrict__ gpow, Fr* __restrict__ out,
                            size_t total) {
    const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= total) return;
    out[k] = k < m ? in[k] * gpow[k] : Fr::zero();
}

// "T Polynomial is not divisible" <=> the gate equation fails on some row of H (the permutation and L_1 parts
// vanish on H by construction of Z, checked in round 2).  flags[0].
struct RowArgs {
    const Fr *A, *B, *C, *QM4, *QL4, *QR4, *QO4, *QC4, *pub;
    uint32_t n, n_pub;
    int* flags;
};
__global__ void __launch_bounds__(256) k_rowcheck(RowArgs q) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= q.n) return;
    const size_t j = (size_t)4 * i;
    const Fr a = q.A[i], b = q.B[i], c = q.C[i];
    Fr v = (a * b) * q.QM4[j] + a * q.QL4[j] + b * q.QR4[j] + c * q.QO4[j] + q.QC4[j];
    if (i < q.n_pub) v = v - q.pub[i];
    if (!v.is_zero()) atomicOr(&q.flags[0], 1);
}

// Quotient on the coset: t(x) = N'(x) / Z_H(x), x = g w_4n^i, with N' the full numerator of the blinded
// polynomials (A.2 round 3: T + Z_H Tz = N' as polynomials, so t is the same polynomial snarkjs obtains from
// its two half-computations).
struct R3Args {
    const Fr *A, *B, *C, *Z;  // blinded a, b, c, z on the coset (4n values each)
    const Fr *QM, *QL, *QR, *QO, *QC, *S1, *S2, *S3, *LAG, *pub, *W4n;
    Fr g, beta, beta_g, gamma, alpha, alpha2, k1, k2;
    Fr zh_inv[4];
    uint32_t n, power, n_pub, k_small;  // k_small: k1 == 2 and k2 == 3
    Fr* T;
};
__global__ void __launch_bounds__(128) k_round3(R3Args q) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t n4 = (size_t)4 * q.n;
    if (i >= n4) return;
    const Fr a = q.A[i], b = q.B[i], c = q.C[i], z = q.Z[i];
    const Fr zw = q.Z[(i + 4) & (n4 - 1)];
    Fr pl = Fr::zero();
    for (uint32_t j = 0; j < q.n_pub; j++) pl = pl - q.LAG[(size_t)j * n4 + i] * q.pub[j];
    const Fr gate = (a * b) * q.QM[i] + a * q.QL[i] + b * q.QR[i] + c * q.QO[i] + q.QC[i] + pl;
    // beta * x with x = g w^i: one product with the host's beta * g; k1 = 2, k2 = 3 (every power snarkjs supports
    // here) turn the two other coset shifts into additions
    const Fr bx = q.beta_g * domain_pow(q.W4n, q.power + 2, i);
    const Fr bx2 = q.k_small ? bx + bx : bx * q.k1;
    const Fr bx3 = q.k_small ? bx2 + bx : bx * q.k2;
    const Fr p1 = ((a + bx + q.gamma) * (b + bx2 + q.gamma)) * ((c + bx3 + q.gamma) * z);
    const Fr p2 = ((a + q.beta * q.S1[i] + q.gamma) * (b + q.beta * q.S2[i] + q.gamma)) *
                  ((c + q.beta * q.S3[i] + q.gamma) * zw);
    const Fr l1 = (z - Fr::one()) * q.LAG[i] * q.alpha2;
    q.T[i] = (gate + (p1 - p2) * q.alpha + l1) * q.zh_inv[i & 3];
}
// coefficients above 3n + 5 must vanish; flags[1]
__global__ void k_check_high(const Fr* __restrict__ t, uint32_t n, int* __restrict__ flags) {
    const size_t i = (size_t)3 * n + 6 + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)4 * n) return;
    if (!t[i].is_zero()) atomicOr(&flags[1], 1);
}

struct R4Args {
    const Fr *pol_z, *qm, *ql, *qr, *qo, *qc, *s3;
    Fr coefz, coef_ab, ea, eb, ec, e3;
    uint32_t n;
    Fr* pol_r;
};
__global__ void k_pol_r(R4Args g) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.n + 3) return;
    Fr v = g.coefz * g.pol_z[i];
    if (i < g.n)
        v = v + g.coef_ab * g.qm[i] + g.ea * g.ql[i] + g.eb * g.qr[i] + g.ec * g.qo[i] + g.qc[i] - g.e3 * g.s3[i];
    g.pol_r[i] = v;
}

struct R5Args {
    const Fr *pol_t, *pol_r, *pol_a, *pol_b, *pol_c, *s1, *s2;
    Fr xim, xi2m, v[7], w0_sub;
    uint32_t n;
    Fr* out;
};
__global__ void k_pol_wxi(R5Args g) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t n = g.n;
    if (i >= n + 6) return;
    Fr w = g.xi2m * g.pol_t[(size_t)2 * n + i];
    if (i < n + 3) w = w + g.v[1] * g.pol_r[i];
    if (i < n + 2) w = w + g.v[2] * g.pol_a[i] + g.v[3] * g.pol_b[i] + g.v[4] * g.pol_c[i];
    if (i < n) w = w + g.pol_t[i] + g.xim * g.pol_t[(size_t)n + i] + g.v[5] * g.s1[i] + g.v[6] * g.s2[i];
    if (i == 0) w = w - g.w0_sub;
    g.out[i] = w;
}
__global__ void k_sub_at0(Fr* p, Fr v) { p[0] = p[0] - v; }

Fr fr_from_le(const uint8_t* p) {
    Fr f;
    memcpy(f.v, p, 32);
    return f;
}
bool fr_is_canonical(const Fr& f) {
    for (int i = 7; i >= 0; i--) {
        if (f.v[i] < FrParams::mod(i)) return true;
        if (f.v[i] > FrParams::mod(i)) return false;
    }
    return false;
}

}  // namespace

// ------------------------------------------------------------------ zkey load
extern "C" void nzcb_zkey_free(nzcb_zkey* zk) {
    if (!zk) return;
    if (zk->ctx) {
        cudaSetDevice(zk->ctx->device);
        cudaStreamSynchronize(zk->ctx->stream);
    }
    for (int i = 0; i < 3; i++) cudaFree(zk->d_map[i]);
    cudaFree(zk->d_add_a);
    cudaFree(zk->d_add_b);
    cudaFree(zk->d_add_out);
    cudaFree(zk->d_add_ac);
    cudaFree(zk->d_add_bc);
    for (int i = 0; i < 5; i++) cudaFree(zk->d_q[i]);
    cudaFree(zk->d_sigma);
    cudaFree(zk->d_lag);
    g1_table_free(&zk->tab);
    g1_table_free(&zk->tab_lag);
    for (int i = 0; i < 8; i++) cudaFree(zk->d_cos[i]);
    cudaFree(zk->d_cos_lag);
    cudaFree(zk->d_gpow);
    cudaFree(zk->d_ginv);
    delete zk;
}

extern "C" int32_t nzcb_zkey_info(const nzcb_zkey* zk, uint32_t* n_vars, uint32_t* n_public, uint32_t* domain_size,
                                  uint32_t* n_additions, uint32_t* n_constraints) {
    if (!zk) return NZCB_E_INVALID;
    if (n_vars) *n_vars = zk->n_vars;
    if (n_public) *n_public = zk->n_public;
    if (domain_size) *domain_size = zk->n;
    if (n_additions) *n_additions = zk->n_add;
    if (n_constraints) *n_constraints = zk->n_cons;
    return 0;
}

#define ZK_CUDA(call)                                                                                      \
    do {                                                                                                   \
        cudaError_t e__ = (call);                                                                          \
        if (e__ != cudaSuccess) {                                                                          \
            ctx->fail(NZCB_E_CUDA, "CUDA error %s at %s:%d", cudaGetErrorString(e__), __FILE__, __LINE__); \
            nzcb_zkey_free(zk);                                                                            \
            return NZCB_E_CUDA;                                                                            \
        }                                                                                                  \
    } while (0)

extern "C" int32_t nzcb_zkey_load(nzcb_ctx* ctx, const uint8_t* data, size_t len, nzcb_zkey** out) {
    if (!ctx || !data || !out) return NZCB_E_INVALID;
    *out = nullptr;
    std::map<uint32_t, Section> sec;
    NZ_TRY(parse_binfile(ctx, data, len, "zkey", sec));
    for (uint32_t id = 1; id <= 14; id++)
        if (!sec.count(id)) return ctx->fail(NZCB_E_INVALID, "zkey file: section %u missing", id);
    uint32_t proto = 0;
    if (sec[1].size < 4) return ctx->fail(NZCB_E_INVALID, "zkey file: bad section 1");
    memcpy(&proto, sec[1].p, 4);
    if (proto != 2) return ctx->fail(NZCB_E_INVALID, "zkey file is not plonk");
    const uint8_t* h = sec[2].p;
    if (sec[2].size < 4 + 32 + 4 + 32 + 20 + 64 + 8 * 64 + 128) return ctx->fail(NZCB_E_INVALID, "zkey file: short header");
    uint32_t n8q, n8r;
    memcpy(&n8q, h, 4);
    memcpy(&n8r, h + 36, 4);
    if (n8q != 32 || n8r != 32 || memcmp(h + 4, Q_LE, 32) != 0 || memcmp(h + 40, R_LE, 32) != 0)
        return ctx->fail(NZCB_E_INVALID, "zkey file: curve is not bn128");
    nzcb_zkey* zk = new nzcb_zkey();
    zk->ctx = ctx;
    uint32_t f[5];
    memcpy(f, h + 72, 20);
    zk->n_vars = f[0];
    zk->n_public = f[1];
    zk->n = f[2];
    zk->n_add = f[3];
    zk->n_cons = f[4];
    memcpy(zk->k1.v, h + 92, 32);
    memcpy(zk->k2.v, h + 124, 32);
    const uint64_t n = zk->n;
    if (n < 8 || (n & (n - 1)) || zk->n_cons > n || zk->n_add > zk->n_vars) {
        delete zk;
        return ctx->fail(NZCB_E_INVALID, "zkey file: inconsistent header (domainSize %u, nConstraints %u)", (unsigned)n,
                         f[4]);
    }
    zk->power = 0;
    while (((uint64_t)1 << zk->power) < n) zk->power++;
    const uint32_t n_lag = zk->n_public > 1 ? zk->n_public : 1;
    bool ok = sec[3].size == (uint64_t)zk->n_add * 72;
    for (int c = 0; c < 3; c++) ok = ok && sec[4 + c].size == (uint64_t)zk->n_cons * 4;
    for (int k = 0; k < 5; k++) ok = ok && sec[7 + k].size == 5 * n * 32;
    ok = ok && sec[12].size == 15 * n * 32 && sec[13].size == (uint64_t)n_lag * 5 * n * 32 && sec[14].size == (n + 6) * 64;
    if (!ok) {
        delete zk;
        return ctx->fail(NZCB_E_INVALID, "zkey file: section sizes do not match the header");
    }
    ZK_CUDA(cudaSetDevice(ctx->device));
    for (int c = 0; c < 3; c++) {
        ZK_CUDA(cudaMalloc(&zk->d_map[c], std::max<size_t>(4, sec[4 + c].size)));
        ZK_CUDA(cudaMemcpyAsync(zk->d_map[c], sec[4 + c].p, sec[4 + c].size, cudaMemcpyHostToDevice, ctx->stream));
    }
    for (int k = 0; k < 5; k++) {
        ZK_CUDA(cudaMalloc(&zk->d_q[k], sec[7 + k].size));
        ZK_CUDA(cudaMemcpyAsync(zk->d_q[k], sec[7 + k].p, sec[7 + k].size, cudaMemcpyHostToDevice, ctx->stream));
    }
    ZK_CUDA(cudaMalloc(&zk->d_sigma, sec[12].size));
    ZK_CUDA(cudaMemcpyAsync(zk->d_sigma, sec[12].p, sec[12].size, cudaMemcpyHostToDevice, ctx->stream));
    ZK_CUDA(cudaMalloc(&zk->d_lag, sec[13].size));
    ZK_CUDA(cudaMemcpyAsync(zk->d_lag, sec[13].p, sec[13].size, cudaMemcpyHostToDevice, ctx->stream));
    {   // SRS: upload section 14 and precompute the window shifts 2^(c w) [tau^i]G1 once per key
        G1Affine* d_ptau = (G1Affine*)ctx->scratch_get("zk_ptau_in", sec[14].size);
        if (!d_ptau) {
            nzcb_zkey_free(zk);
            return ctx->fail(NZCB_E_NOMEM, "zkey: cannot allocate the SRS staging buffer");
        }
        ZK_CUDA(cudaMemcpyAsync(d_ptau, sec[14].p, sec[14].size, cudaMemcpyHostToDevice, ctx->stream));
        int rc = g1_table_build(ctx, d_ptau, n + 6, &zk->tab);
        G1Affine* d_lagpts = (G1Affine*)ctx->scratch_get("zk_lag_pts", (n + 4) * sizeof(G1Affine));
        if (rc == 0 && !d_lagpts) rc = ctx->fail(NZCB_E_NOMEM, "zkey: cannot allocate the Lagrange-basis staging buffer");
        if (rc == 0) rc = g1_lagrange_basis(ctx, d_ptau, zk->power, d_lagpts);
        if (rc == 0) {
            const size_t src[4] = {0, 1, (size_t)n, (size_t)n + 1};
            for (int k = 0; k < 4 && rc == 0; k++)
                if (cudaMemcpyAsync(d_lagpts + n + k, d_ptau + src[k], sizeof(G1Affine), cudaMemcpyDeviceToDevice,
                                    ctx->stream) != cudaSuccess)
                    rc = ctx->fail(NZCB_E_CUDA, "zkey: copy of the blinding bases failed");
        }
        // wire values are mostly 0 / +-1 / bytes: few digits whatever the window, so a 14-bit window (8 K buckets
        // per commitment instead of 512 K) keeps the bucket reduction from dominating round 1
        if (rc == 0) {
            uint32_t c_lag = zk->power > 15 ? 14 : 0;
            const char* env = getenv("NZCB_MSM_LAGRANGE_WINDOW");
            if (env && atoi(env) >= 4 && atoi(env) <= 20) c_lag = (uint32_t)atoi(env);
            rc = g1_table_build(ctx, d_lagpts, n + 4, &zk->tab_lag, c_lag);
        }
        if (rc != 0) {
            nzcb_zkey_free(zk);
            return rc;
        }
    }

    {   // coset evaluations for round 3
        zk->g = Fr::from_u64(7);
        Fr gn = zk->g;
        for (uint32_t i = 0; i < zk->power; i++) gn = gn.sqr();
        const Fr w4 = fr_root_host(2);
        Fr wp = Fr::one();
        bool bad = false;
        for (int pp = 0; pp < 4; pp++) {
            const Fr d = gn * wp - Fr::one();
            bad = bad || d.is_zero();
            zk->zh_inv[pp] = d.inv();
            wp = wp * w4;
        }
        if (bad) {
            nzcb_zkey_free(zk);
            return ctx->fail(NZCB_E_INVALID, "zkey: the coset shift lies in the 4n domain");
        }
        const size_t n4 = 4 * n;
        ZK_CUDA(cudaMalloc(&zk->d_gpow, (n + 8) * sizeof(Fr)));
        ZK_CUDA(cudaMalloc(&zk->d_ginv, n4 * sizeof(Fr)));
        k_pow_table<<<div_up(n + 8, 256), 256, 0, ctx->stream>>>(zk->d_gpow, zk->g, Fr::one(), n + 8);
        k_pow_table<<<div_up(n4, 256), 256, 0, ctx->stream>>>(zk->d_ginv, zk->g.inv(), Fr::from_u64(n4).inv(), n4);
        ctx->launches += 2;
        auto to_coset = [&](const Fr* coef, Fr* out) -> int {
            k_scale_pad<<<div_up(n4, 256), 256, 0, ctx->stream>>>(coef, n, zk->d_gpow, out, n4);
            ctx->launches++;
            return ntt_dev(ctx, out, zk->power + 2, false);
        };
        for (int k = 0; k < 8; k++) {
            ZK_CUDA(cudaMalloc(&zk->d_cos[k], n4 * sizeof(Fr)));
            const Fr* coef = k < 5 ? zk->d_q[k] : zk->d_sigma + (size_t)(k - 5) * 5 * n;
            if (to_coset(coef, zk->d_cos[k]) != 0) {
                nzcb_zkey_free(zk);
                return NZCB_E_CUDA;
            }
        }
        ZK_CUDA(cudaMalloc(&zk->d_cos_lag, (size_t)n_lag * n4 * sizeof(Fr)));
        for (uint32_t j = 0; j < n_lag; j++) {
            if (to_coset(zk->d_lag + (size_t)j * 5 * n, zk->d_cos_lag + (size_t)j * n4) != 0) {
                nzcb_zkey_free(zk);
                return NZCB_E_CUDA;
            }
        }
        ZK_CUDA(cudaGetLastError());
    }

    // additions: level-schedule.  level(i) = 1 + max(level of operands that are themselves additions)
    const uint32_t n_w = zk->n_vars - zk->n_add;
    const uint32_t na = zk->n_add;
    std::vector<uint32_t> lvl(na), ia(na), ib(na);
    uint32_t max_lvl = 0;
    const uint8_t* ap = sec[3].p;
    for (uint32_t i = 0; i < na; i++) {
        memcpy(&ia[i], ap + (size_t)i * 72, 4);
        memcpy(&ib[i], ap + (size_t)i * 72 + 4, 4);
        uint32_t l = 0;
        for (uint32_t s : {ia[i], ib[i]}) {
            if (s >= n_w && s < zk->n_vars) {
                const uint32_t j = s - n_w;
                if (j >= i) {
                    nzcb_zkey_free(zk);
                    return ctx->fail(NZCB_E_INVALID, "zkey file: addition %u depends on a later addition", i);
                }
                l = std::max(l, lvl[j] + 1);
            }
        }
        lvl[i] = l;
        max_lvl = std::max(max_lvl, l);
    }
    std::vector<uint32_t> cnt(max_lvl + 2, 0);
    for (uint32_t i = 0; i < na; i++) cnt[lvl[i] + 1]++;
    for (uint32_t l = 0; l <= max_lvl; l++) cnt[l + 1] += cnt[l];
    zk->level_off.assign(cnt.begin(), cnt.end());
    if (na == 0) zk->level_off.assign(1, 0);
    std::vector<uint32_t> pos(cnt.begin(), cnt.end() - 1), sa(na), sb(na), so(na);
    std::vector<Fr> sac(na), sbc(na);
    for (uint32_t i = 0; i < na; i++) {
        const uint32_t d = pos[lvl[i]]++;
        sa[d] = ia[i];
        sb[d] = ib[i];
        so[d] = i;
        memcpy(sac[d].v, ap + (size_t)i * 72 + 8, 32);
        memcpy(sbc[d].v, ap + (size_t)i * 72 + 40, 32);
    }
    const size_t na1 = std::max<uint32_t>(na, 1);
    ZK_CUDA(cudaMalloc(&zk->d_add_a, na1 * 4));
    ZK_CUDA(cudaMalloc(&zk->d_add_b, na1 * 4));
    ZK_CUDA(cudaMalloc(&zk->d_add_out, na1 * 4));
    ZK_CUDA(cudaMalloc(&zk->d_add_ac, na1 * 32));
    ZK_CUDA(cudaMalloc(&zk->d_add_bc, na1 * 32));
    if (na) {
        ZK_CUDA(cudaMemcpyAsync(zk->d_add_a, sa.data(), (size_t)na * 4, cudaMemcpyHostToDevice, ctx->stream));
        ZK_CUDA(cudaMemcpyAsync(zk->d_add_b, sb.data(), (size_t)na * 4, cudaMemcpyHostToDevice, ctx->stream));
        ZK_CUDA(cudaMemcpyAsync(zk->d_add_out, so.data(), (size_t)na * 4, cudaMemcpyHostToDevice, ctx->stream));
        ZK_CUDA(cudaMemcpyAsync(zk->d_add_ac, sac.data(), (size_t)na * 32, cudaMemcpyHostToDevice, ctx->stream));
        ZK_CUDA(cudaMemcpyAsync(zk->d_add_bc, sbc.data(), (size_t)na * 32, cudaMemcpyHostToDevice, ctx->stream));
    }
    ZK_CUDA(cudaStreamSynchronize(ctx->stream));
    *out = zk;
    return 0;
}

// ------------------------------------------------------------------ prove
namespace {

struct Bufs {
    Fr *W, *A, *B, *C, *pol_a, *pol_b, *pol_c, *pol_z, *A4, *B4, *C4, *Z4, *num, *den, *T, *pol_r, *pol_wxi,
        *quot, *quot2, *vals, *pub;
    G1XYZZ* pts;
    int* flags;
};

#define GETBUF(field, name, count)                                                            \
    do {                                                                                      \
        b.field = (decltype(b.field))ctx->scratch_get(name, (size_t)(count) * sizeof(*b.field)); \
        if (!b.field) return ctx->fail(NZCB_E_NOMEM, "prove: cannot allocate " name);         \
    } while (0)

int get_bufs(nzcb_ctx* ctx, const nzcb_zkey* zk, Bufs& b) {
    const size_t n = zk->n;
    GETBUF(W, "pv_W", zk->n_vars + 1);
    GETBUF(A, "pv_A", n + 8);
    GETBUF(B, "pv_B", n + 8);
    GETBUF(C, "pv_C", n + 8);
    GETBUF(pol_a, "pv_pol_a", n + 8);
    GETBUF(pol_b, "pv_pol_b", n + 8);
    GETBUF(pol_c, "pv_pol_c", n + 8);
    GETBUF(pol_z, "pv_pol_z", n + 8);
    GETBUF(A4, "pv_A4", 4 * n);
    GETBUF(B4, "pv_B4", 4 * n);
    GETBUF(C4, "pv_C4", 4 * n);
    GETBUF(Z4, "pv_Z4", 4 * n);
    GETBUF(num, "pv_num", n);
    GETBUF(den, "pv_den", n);
    GETBUF(T, "pv_T", 4 * n);
    GETBUF(pol_r, "pv_pol_r", n + 8);
    GETBUF(pol_wxi, "pv_pol_wxi", n + 8);
    GETBUF(quot, "pv_quot", n + 8);
    GETBUF(quot2, "pv_quot2", n + 8);
    GETBUF(vals, "pv_vals", 16);
    GETBUF(pub, "pv_pub", zk->n_public + 1);
    GETBUF(pts, "pv_pts", 4);
    GETBUF(flags, "pv_flags", 4);
    return 0;
}

// NZCB_TRACE=1: per-phase device times (CUDA events on the ctx stream) to stderr
struct Tracer {
    nzcb_ctx* ctx;
    bool on;
    std::vector<std::pair<const char*, cudaEvent_t>> marks;
    explicit Tracer(nzcb_ctx* c) : ctx(c) {
        const char* e = getenv("NZCB_TRACE");
        on = e && e[0] == '1';
        mark("start");
    }
    void mark(const char* name) {
        if (!on) return;
        cudaEvent_t ev;
        cudaEventCreate(&ev);
        cudaEventRecord(ev, ctx->stream);
        marks.push_back({name, ev});
    }
    ~Tracer() {
        if (!on) return;
        cudaStreamSynchronize(ctx->stream);
        for (size_t i = 1; i < marks.size(); i++) {
            float ms = 0;
            cudaEventElapsedTime(&ms, marks[i - 1].second, marks[i].second);
            fprintf(stderr, "[nzcb trace] %-14s %9.3f ms\n", marks[i].first, ms);
        }
        float tot = 0;
        if (marks.size() > 1) cudaEventElapsedTime(&tot, marks.front().second, marks.back().second);
        fprintf(stderr, "[nzcb trace] %-14s %9.3f ms\n", "total", tot);
        for (auto& m : marks) cudaEventDestroy(m.second);
    }
};

// evaluations -> blinded coefficient polynomial   [first half of the role of snarkjs to4T]
int to_coef(nzcb_ctx* ctx, const nzcb_zkey* zk, const Fr* d_evals, Fr* d_pol, const Fr* pz, int k) {
    const size_t n = zk->n;
    NZ_CUDA(ctx, cudaMemcpyAsync(d_pol, d_evals, n * sizeof(Fr), cudaMemcpyDeviceToDevice, ctx->stream));
    NZ_TRY(ntt_dev(ctx, d_pol, zk->power, true));
    Blind bl;
    bl.k = k;
    for (int i = 0; i < 3; i++) bl.pz[i] = i < k ? pz[i] : Fr::zero();
    NZ_LAUNCH(ctx, k_blind, 1, 32, 0, d_pol, n, bl);
    return 0;
}
// blinded polynomial (n + k coefficients) -> its 4n evaluations on the coset g * H_4n (round 3)
int to_coset(nzcb_ctx* ctx, const nzcb_zkey* zk, const Fr* d_pol, Fr* d_ext, int k) {
    const size_t n = zk->n;
    NZ_LAUNCH(ctx, k_scale_pad, div_up(4 * n, 256), 256, 0, d_pol, n + (size_t)k, zk->d_gpow, d_ext, 4 * n);
    return ntt_dev(ctx, d_ext, zk->power + 2, false);
}

// While alive, everything enqueued through the ctx goes to its side stream, which first waits for what the main stream
// holds so far (fork = true) -- the commitments of a round run beside its transforms.  The destructor switches back;
// whoever needs the results synchronises the side stream (msm_table_finish does).
struct SideStream {
    nzcb_ctx* ctx;
    cudaStream_t main;
    bool ok = false;
    SideStream(nzcb_ctx* c, bool fork) : ctx(c), main(c->stream) {
        if (!ctx->side && cudaStreamCreateWithFlags(&ctx->side, cudaStreamNonBlocking) != cudaSuccess) return;
        if (!ctx->ev_fork && cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming) != cudaSuccess) return;
        if (fork && (cudaEventRecord(ctx->ev_fork, main) != cudaSuccess ||
                     cudaStreamWaitEvent(ctx->side, ctx->ev_fork, 0) != cudaSuccess))
            return;
        ctx->stream = ctx->side;
        ok = true;
    }
    ~SideStream() { ctx->stream = main; }
};

// d_w_le: device, n_w canonical little-endian witness values (NOT yet Montgomery); may alias b.w_le.
// h_pub_le: host copy of w[1..nPublic] (canonical LE) for the transcript.
int prove_core(nzcb_ctx* ctx, const nzcb_zkey* zk, const Fr* d_w_le, const uint8_t* h_pub_le, const uint8_t* blinders_le,
               nzcb_proof* out, uint8_t* public_le);

int prove_one(nzcb_ctx* ctx, const nzcb_zkey* zk, const uint8_t* wtns, size_t wtns_len, const uint8_t* blinders_le,
              nzcb_proof* out, uint8_t* public_le) {
    NZ_CUDA(ctx, cudaSetDevice(ctx->device));
    // ---- step 0: parse + check the witness file
    std::map<uint32_t, Section> ws;
    NZ_TRY(parse_binfile(ctx, wtns, wtns_len, "wtns", ws));
    if (!ws.count(1) || !ws.count(2) || ws[1].size < 40) return ctx->fail(NZCB_E_INVALID, "wtns file: missing sections");
    uint32_t n8, n_wit;
    memcpy(&n8, ws[1].p, 4);
    memcpy(&n_wit, ws[1].p + 36, 4);
    if (n8 != 32 || memcmp(ws[1].p + 4, R_LE, 32) != 0)
        return ctx->fail(NZCB_E_WITNESS, "Curve of the witness does not match the curve of the proving key");
    const uint32_t n_w = zk->n_vars - zk->n_add;
    if (n_wit != n_w || ws[2].size != (uint64_t)n_wit * 32)
        return ctx->fail(NZCB_E_WITNESS, "Invalid witness length. Circuit: %u, witness: %u, %u", zk->n_vars, n_wit,
                         zk->n_add);
    if (zk->n_public + 1 > n_w) return ctx->fail(NZCB_E_WITNESS, "Invalid witness length: fewer values than public signals");
    const uint8_t* wv = ws[2].p;
    Fr* d_w = (Fr*)ctx->scratch_get("pv_w_le", (size_t)zk->n_vars * sizeof(Fr));
    if (!d_w) return ctx->fail(NZCB_E_NOMEM, "prove: cannot allocate the witness buffer");
    NZ_CUDA(ctx, cudaMemcpyAsync(d_w, wv, (size_t)n_w * 32, cudaMemcpyHostToDevice, ctx->stream));
    return prove_core(ctx, zk, d_w, wv + 32, blinders_le, out, public_le);
}

int prove_core(nzcb_ctx* ctx, const nzcb_zkey* zk, const Fr* d_w_le, const uint8_t* h_pub_le, const uint8_t* blinders_le,
               nzcb_proof* out, uint8_t* public_le) {
    const uint32_t n_w = zk->n_vars - zk->n_add;
    const uint32_t n = zk->n, n_pub = zk->n_public;

    // blinders b1..b9 (Montgomery); index 0 unused.  Without injected values: one read of the OS CSPRNG for all nine
    // (254-bit candidates, rejected above r -- probability 1/4 each -- and redrawn from the same descriptor)
    Fr bl[10];
    bl[0] = Fr::zero();
    FILE* rnd = nullptr;
    if (!blinders_le) {
        rnd = fopen("/dev/urandom", "rb");
        if (!rnd) return ctx->fail(NZCB_E_INVALID, "cannot open the OS CSPRNG");
    }
    for (int i = 1; i <= 9; i++) {
        Fr v;
        if (blinders_le) {
            v = fr_from_le(blinders_le + (i - 1) * 32);
            if (!fr_is_canonical(v)) return ctx->fail(NZCB_E_INVALID, "blinder b%d is not a canonical Fr element", i);
        } else {
            do {
                if (fread(v.v, 1, 32, rnd) != 32) {
                    fclose(rnd);
                    return ctx->fail(NZCB_E_INVALID, "cannot read the OS CSPRNG");
                }
                v.v[7] &= 0x3fffffffu;
            } while (!fr_is_canonical(v));
        }
        bl[i] = v.to_mont();
    }
    if (rnd) fclose(rnd);

    Bufs b;
    NZ_TRY(get_bufs(ctx, zk, b));
    NZ_CUDA(ctx, cudaMemsetAsync(b.flags, 0, 4 * sizeof(int), ctx->stream));
    NZ_LAUNCH(ctx, k_wtns_to_mont, div_up(n_w, 256), 256, 0, d_w_le, b.W, (size_t)n_w);
    Tracer tr_(ctx);
    for (size_t l = 0; l + 1 < zk->level_off.size(); l++) {
        const uint32_t lo = zk->level_off[l], hi = zk->level_off[l + 1];
        if (hi > lo)
            NZ_LAUNCH(ctx, k_additions, div_up(hi - lo, 256), 256, 0, zk->d_add_a, zk->d_add_b, zk->d_add_ac,
                      zk->d_add_bc, zk->d_add_out, lo, hi, zk->n_vars, n_w, b.W);
    }
    tr_.mark("wtns+additions");
    // ---- round 1
    NZ_LAUNCH(ctx, k_gather, div_up(n, 256), 256, 0, zk->d_map[0], b.W, zk->n_vars, zk->n_cons, n, b.A);
    NZ_LAUNCH(ctx, k_gather, div_up(n, 256), 256, 0, zk->d_map[1], b.W, zk->n_vars, zk->n_cons, n, b.B);
    NZ_LAUNCH(ctx, k_gather, div_up(n, 256), 256, 0, zk->d_map[2], b.W, zk->n_vars, zk->n_cons, n, b.C);
    G1Affine cA, cB, cC, cZ, cT1, cT2, cT3, cWxi, cWxiw;
    {   // Lagrange basis: the scalars are the wire values (mostly 0 / +-1 / bytes) plus four blinding terms.  The
        // commitments only need the evaluations, so they run on the side stream beside the round's transforms.
        NZ_LAUNCH(ctx, k_blind_scalars, 1, 32, 0, b.A, (size_t)n, bl[2], bl[1]);
        NZ_LAUNCH(ctx, k_blind_scalars, 1, 32, 0, b.B, (size_t)n, bl[4], bl[3]);
        NZ_LAUNCH(ctx, k_blind_scalars, 1, 32, 0, b.C, (size_t)n, bl[6], bl[5]);
        const uint32_t* sc[3] = {(const uint32_t*)b.A, (const uint32_t*)b.B, (const uint32_t*)b.C};
        const size_t sn[3] = {(size_t)n + 4, (size_t)n + 4, (size_t)n + 4};
        {
            SideStream side(ctx, true);
            if (!side.ok) return ctx->fail(NZCB_E_CUDA, "prove: cannot set up the side stream");
            NZ_TRY(msm_table_dev(ctx, zk->tab_lag, sc, sn, 3, true, b.pts, true));
        }
        const Fr pa[2] = {bl[2], bl[1]}, pb[2] = {bl[4], bl[3]}, pc[2] = {bl[6], bl[5]};
        NZ_TRY(to_coef(ctx, zk, b.A, b.pol_a, pa, 2));
        NZ_TRY(to_coset(ctx, zk, b.pol_a, b.A4, 2));
        NZ_TRY(to_coef(ctx, zk, b.B, b.pol_b, pb, 2));
        NZ_TRY(to_coset(ctx, zk, b.pol_b, b.B4, 2));
        NZ_TRY(to_coef(ctx, zk, b.C, b.pol_c, pc, 2));
        NZ_TRY(to_coset(ctx, zk, b.pol_c, b.C4, 2));
        tr_.mark("r1 ntt");
        G1Affine r[3];
        {
            SideStream side(ctx, false);
            if (!side.ok) return ctx->fail(NZCB_E_CUDA, "prove: cannot set up the side stream");
            NZ_TRY(msm_table_finish(ctx, b.pts, r, 3));
        }
        cA = r[0]; cB = r[1]; cC = r[2];
    }
    g1_to_be(cA, out->A);
    g1_to_be(cB, out->B);
    g1_to_be(cC, out->C);

    tr_.mark("r1 msm x3");
    // ---- round 2
    std::vector<uint8_t> tr;
    std::vector<Fr> pub_m(n_pub + 1);
    for (uint32_t i = 0; i < n_pub; i++) {
        const Fr v = fr_from_le(h_pub_le + (size_t)i * 32);  // A[i] = w[i+1] for the public-input gates
        if (!fr_is_canonical(v)) return ctx->fail(NZCB_E_WITNESS, "witness value %u is not reduced", i + 1);
        pub_m[i] = v.to_mont();
        uint8_t be[32];
        to_be_bytes(pub_m[i], be);
        append(tr, be, 32);
        if (public_le) memcpy(public_le + (size_t)i * 32, h_pub_le + (size_t)i * 32, 32);
    }
    append(tr, out->A, 64);
    append(tr, out->B, 64);
    append(tr, out->C, 64);
    const Fr beta = hash_to_fr(tr);
    uint8_t be32[32];
    to_be_bytes(beta, be32);
    const Fr gamma = hash_to_fr(std::vector<uint8_t>(be32, be32 + 32));

    const Fr *Wn = nullptr, *W4n = nullptr;
    NZ_TRY(get_twiddles_pub(ctx, zk->power, false, &Wn));
    NZ_TRY(get_twiddles_pub(ctx, zk->power + 2, false, &W4n));
    const size_t N = n;
    const Fr* S1 = zk->d_sigma;
    const Fr* S14 = zk->d_sigma + N;
    const Fr* S2 = zk->d_sigma + 5 * N;
    const Fr* S24 = zk->d_sigma + 6 * N;
    const Fr* S3 = zk->d_sigma + 10 * N;
    const Fr* S34 = zk->d_sigma + 11 * N;
    {
        R2Args a;
        a.A = b.A; a.B = b.B; a.C = b.C;
        a.S14 = S14; a.S24 = S24; a.S34 = S34; a.Wn = Wn;
        a.beta = beta; a.gamma = gamma; a.k1 = zk->k1; a.k2 = zk->k2;
        a.n = n; a.power = zk->power; a.num = b.num; a.den = b.den;
        NZ_LAUNCH(ctx, k_round2_terms, div_up(n, 128), 128, 0, a);
    }
    NZ_TRY(batch_inverse(ctx, b.den, n));
    NZ_LAUNCH(ctx, k_mul_inplace, div_up(n, 256), 256, 0, b.num, b.den, (size_t)n);
    // Z[i] = prod_{j<i} num_j/den_j ; the wrap-around product must be 1
    NZ_TRY(prefix_product(ctx, b.num, n, b.den, b.vals + 0));
    {
        Fr total;
        NZ_CUDA(ctx, cudaMemcpyAsync(&total, b.vals + 0, sizeof(Fr), cudaMemcpyDeviceToHost, ctx->stream));
        NZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (total != Fr::one()) return ctx->fail(NZCB_E_COPY, "Copy constraints does not match");
    }
    tr_.mark("r2 grandprod");
    {   // the commitment needs the coefficients, not the coset evaluations: it starts as soon as they exist
        const Fr pz[3] = {bl[9], bl[8], bl[7]};
        NZ_TRY(to_coef(ctx, zk, b.den, b.pol_z, pz, 3));
        const uint32_t* sc[1] = {(const uint32_t*)b.pol_z};
        const size_t sn[1] = {(size_t)n + 3};
        {
            SideStream side(ctx, true);
            if (!side.ok) return ctx->fail(NZCB_E_CUDA, "prove: cannot set up the side stream");
            NZ_TRY(msm_table_dev(ctx, zk->tab, sc, sn, 1, true, b.pts));
        }
        NZ_TRY(to_coset(ctx, zk, b.pol_z, b.Z4, 3));
        tr_.mark("r2 ntt");
        {
            SideStream side(ctx, false);
            if (!side.ok) return ctx->fail(NZCB_E_CUDA, "prove: cannot set up the side stream");
            NZ_TRY(msm_table_finish(ctx, b.pts, &cZ, 1));
        }
    }
    g1_to_be(cZ, out->Z);

    tr_.mark("r2 msm");
    // ---- round 3
    const Fr alpha = hash_to_fr(std::vector<uint8_t>(out->Z, out->Z + 64));
    const Fr alpha2 = alpha * alpha;
    NZ_CUDA(ctx, cudaMemcpyAsync(b.pub, pub_m.data(), (size_t)(n_pub + 1) * sizeof(Fr), cudaMemcpyHostToDevice, ctx->stream));
    {
        RowArgs q;
        q.A = b.A; q.B = b.B; q.C = b.C;
        q.QM4 = zk->d_q[0] + N; q.QL4 = zk->d_q[1] + N; q.QR4 = zk->d_q[2] + N; q.QO4 = zk->d_q[3] + N; q.QC4 = zk->d_q[4] + N;
        q.pub = b.pub; q.n = n; q.n_pub = n_pub; q.flags = b.flags;
        NZ_LAUNCH(ctx, k_rowcheck, div_up(n, 256), 256, 0, q);
    }
    {
        R3Args g;
        g.A = b.A4; g.B = b.B4; g.C = b.C4; g.Z = b.Z4;
        g.QM = zk->d_cos[0]; g.QL = zk->d_cos[1]; g.QR = zk->d_cos[2]; g.QO = zk->d_cos[3]; g.QC = zk->d_cos[4];
        g.S1 = zk->d_cos[5]; g.S2 = zk->d_cos[6]; g.S3 = zk->d_cos[7]; g.LAG = zk->d_cos_lag; g.pub = b.pub; g.W4n = W4n;
        g.g = zk->g; g.beta = beta; g.gamma = gamma; g.alpha = alpha; g.alpha2 = alpha2; g.k1 = zk->k1; g.k2 = zk->k2;
        g.beta_g = beta * zk->g;
        g.k_small = (zk->k1 == Fr::from_u64(2) && zk->k2 == Fr::from_u64(3)) ? 1u : 0u;
        for (int i = 0; i < 4; i++) g.zh_inv[i] = zk->zh_inv[i];
        g.n = n; g.power = zk->power; g.n_pub = n_pub; g.T = b.T;
        NZ_LAUNCH(ctx, k_round3, div_up(4 * N, 128), 128, 0, g);
    }
    tr_.mark("r3 quotient");
    NZ_TRY(ntt_dev_tab(ctx, b.T, zk->power + 2, true, zk->d_ginv));  // t(gX) -> t(X): factor g^-k / 4n per coefficient
    NZ_LAUNCH(ctx, k_check_high, div_up(n, 256), 256, 0, b.T, n, b.flags);
    {
        int fl[4];
        NZ_CUDA(ctx, cudaMemcpyAsync(fl, b.flags, sizeof(fl), cudaMemcpyDeviceToHost, ctx->stream));
        NZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (fl[0] || fl[1]) return ctx->fail(NZCB_E_DIVIDE, "T Polynomial is not divisible");
    }
    tr_.mark("r3 intt");
    Fr* pol_t = b.T;  // 3n + 6 coefficients
    {
        const uint32_t* sc[3] = {(const uint32_t*)pol_t, (const uint32_t*)(pol_t + N), (const uint32_t*)(pol_t + 2 * N)};
        const size_t sn[3] = {N, N, N + 6};
        G1Affine r[3];
        NZ_TRY(msm_table_dev(ctx, zk->tab, sc, sn, 3, true, b.pts));
        NZ_TRY(msm_table_finish(ctx, b.pts, r, 3));
        cT1 = r[0]; cT2 = r[1]; cT3 = r[2];
    }
    g1_to_be(cT1, out->T1);
    g1_to_be(cT2, out->T2);
    g1_to_be(cT3, out->T3);

    tr_.mark("r3 msm x3");
    // ---- round 4
    tr.clear();
    append(tr, out->T1, 64);
    append(tr, out->T2, 64);
    append(tr, out->T3, 64);
    const Fr xi = hash_to_fr(tr);
    const Fr wn = fr_root_host(zk->power);
    const Fr xiw = xi * wn;
    // vals: 1 eval_a, 2 eval_b, 3 eval_c, 4 eval_s1, 5 eval_s2, 6 eval_t, 7 eval_zw, 8 eval_r, 9/10 remainders
    {   // the seven evaluations of round 4 in one chain of launches (they are independent)
        const Fr* ps[7] = {b.pol_a, b.pol_b, b.pol_c, S1, S2, pol_t, b.pol_z};
        const size_t ns[7] = {N + 2, N + 2, N + 2, N, N, 3 * N + 6, N + 3};
        const Fr at[7] = {xi, xi, xi, xi, xi, xi, xiw};
        Fr* vs[7] = {b.vals + 1, b.vals + 2, b.vals + 3, b.vals + 4, b.vals + 5, b.vals + 6, b.vals + 7};
        NZ_TRY(poly_horner_multi(ctx, 7, ps, ns, at, vs));
    }
    Fr ev[16];
    NZ_CUDA(ctx, cudaMemcpyAsync(ev, b.vals, 8 * sizeof(Fr), cudaMemcpyDeviceToHost, ctx->stream));
    NZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const Fr ea = ev[1], eb = ev[2], ec = ev[3], es1 = ev[4], es2 = ev[5], et = ev[6], ezw = ev[7];
    const Fr coef_ab = ea * eb;
    const Fr bxi = beta * xi;
    const Fr e2 = (ea + bxi + gamma) * (eb + bxi * zk->k1 + gamma) * (ec + bxi * zk->k2 + gamma) * alpha;
    const Fr e3 = (ea + beta * es1 + gamma) * (eb + beta * es2 + gamma) * beta * ezw * alpha;
    Fr xim = xi;
    for (uint32_t i = 0; i < zk->power; i++) xim = xim.sqr();
    const Fr eval_l1 = (xim - Fr::one()) * ((xi - Fr::one()) * Fr::from_u64(n)).inv();
    const Fr e4 = eval_l1 * alpha2;
    {
        R4Args g;
        g.pol_z = b.pol_z; g.qm = zk->d_q[0]; g.ql = zk->d_q[1]; g.qr = zk->d_q[2]; g.qo = zk->d_q[3]; g.qc = zk->d_q[4];
        g.s3 = S3; g.coefz = e2 + e4; g.coef_ab = coef_ab; g.ea = ea; g.eb = eb; g.ec = ec; g.e3 = e3;
        g.n = n; g.pol_r = b.pol_r;
        NZ_LAUNCH(ctx, k_pol_r, div_up(n + 3, 256), 256, 0, g);
    }
    NZ_TRY(poly_horner(ctx, b.pol_r, N + 3, xi, b.vals + 8, nullptr));
    Fr er;
    NZ_CUDA(ctx, cudaMemcpyAsync(&er, b.vals + 8, sizeof(Fr), cudaMemcpyDeviceToHost, ctx->stream));
    NZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    to_be_bytes(ea, out->eval_a);
    to_be_bytes(eb, out->eval_b);
    to_be_bytes(ec, out->eval_c);
    to_be_bytes(es1, out->eval_s1);
    to_be_bytes(es2, out->eval_s2);
    to_be_bytes(ezw, out->eval_zw);
    to_be_bytes(er, out->eval_r);

    tr_.mark("r4 evals");
    // ---- round 5
    tr.clear();
    append(tr, out->eval_a, 32);
    append(tr, out->eval_b, 32);
    append(tr, out->eval_c, 32);
    append(tr, out->eval_s1, 32);
    append(tr, out->eval_s2, 32);
    append(tr, out->eval_zw, 32);
    append(tr, out->eval_r, 32);
    Fr v[7];
    v[0] = Fr::one();
    v[1] = hash_to_fr(tr);
    for (int i = 2; i <= 6; i++) v[i] = v[i - 1] * v[1];
    {
        R5Args g;
        g.pol_t = pol_t; g.pol_r = b.pol_r; g.pol_a = b.pol_a; g.pol_b = b.pol_b; g.pol_c = b.pol_c; g.s1 = S1; g.s2 = S2;
        g.xim = xim; g.xi2m = xim * xim;
        for (int i = 0; i < 7; i++) g.v[i] = v[i];
        g.w0_sub = et + v[1] * er + v[2] * ea + v[3] * eb + v[4] * ec + v[5] * es1 + v[6] * es2;
        g.n = n; g.out = b.pol_wxi;
        NZ_LAUNCH(ctx, k_pol_wxi, div_up(n + 6, 256), 256, 0, g);
    }
    NZ_TRY(poly_horner(ctx, b.pol_wxi, N + 6, xi, b.vals + 9, b.quot));
    tr_.mark("r5 wxi poly");
    // W_{xi w} = (pol_z - eval_zw) / (X - xi w)
    NZ_CUDA(ctx, cudaMemcpyAsync(b.pol_wxi, b.pol_z, (N + 3) * sizeof(Fr), cudaMemcpyDeviceToDevice, ctx->stream));
    NZ_LAUNCH(ctx, k_sub_at0, 1, 1, 0, b.pol_wxi, ezw);
    NZ_TRY(poly_horner(ctx, b.pol_wxi, N + 3, xiw, b.vals + 10, b.quot2));
    {
        const uint32_t* sc[2] = {(const uint32_t*)b.quot, (const uint32_t*)b.quot2};
        const size_t sn[2] = {N + 6, N + 3};
        G1Affine r[2];
        NZ_TRY(msm_table_dev(ctx, zk->tab, sc, sn, 2, true, b.pts));
        NZ_TRY(msm_table_finish(ctx, b.pts, r, 2));
        cWxi = r[0]; cWxiw = r[1];
    }
    tr_.mark("r5 rest");
    Fr rem[2];
    NZ_CUDA(ctx, cudaMemcpyAsync(rem, b.vals + 9, 2 * sizeof(Fr), cudaMemcpyDeviceToHost, ctx->stream));
    NZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (!rem[0].is_zero() || !rem[1].is_zero()) return ctx->fail(NZCB_E_DIVIDE, "Polinomial does not divide");
    g1_to_be(cWxi, out->Wxi);
    g1_to_be(cWxiw, out->Wxiw);
    return 0;
}

}  // namespace

extern "C" int32_t nzcb_plonk_prove(nzcb_ctx* ctx, const nzcb_zkey* zk, const uint8_t* wtns, size_t wtns_len,
                                    const uint8_t* blinders_le, nzcb_proof* out, uint8_t* public_le) {
    if (!ctx || !zk || !wtns || !out) return NZCB_E_INVALID;
    if (zk->ctx != ctx) return ctx->fail(NZCB_E_INVALID, "zkey was loaded on a different context");
    cudaEventRecord(ctx->ev0, ctx->stream);
    const int rc = prove_one(ctx, zk, wtns, wtns_len, blinders_le, out, public_le);
    if (rc != 0) {
        cudaStreamSynchronize(ctx->stream);
        return rc;
    }
    cudaEventRecord(ctx->ev1, ctx->stream);
    NZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    cudaEventElapsedTime(&ctx->last_ms, ctx->ev0, ctx->ev1);
    return 0;
}

// ---- batches: independent proofs on `lanes` of the ctx, one host thread + stream each ------------------------
// witness.cu
namespace nzcb {
int witness_dev(nzcb_ctx* ctx, const nzcb_circuit* c, const Fr* d_inputs, size_t B, Fr* d_wires, int32_t* d_status);
uint32_t circuit_n_total(const nzcb_circuit* c);
uint32_t circuit_n_witness(const nzcb_circuit* c);
uint32_t circuit_n_in(const nzcb_circuit* c);
}  // namespace nzcb

namespace {

int lane_count(size_t B) {
    int L = 4;
    const char* env = getenv("NZCB_LANES");
    if (env && atoi(env) >= 1 && atoi(env) <= 8) L = atoi(env);
    if ((size_t)L > B) L = (int)B;
    return L < 1 ? 1 : L;
}

// Runs item(lane_ctx, g) for g < B on lane_count(B) lanes.  item returns 0, a per-proof error (recorded by the
// item itself) or a fatal one (NZCB_E_CUDA / NZCB_E_NOMEM), which stops the batch.  Device time of the whole batch
// (root-stream events bracketing every lane) goes to ctx->last_ms.
template <class Item, class Prologue>
int run_on_lanes(nzcb_ctx* ctx, const nzcb_zkey* zk, size_t B, Item item, Prologue prologue) {
    NZ_CUDA(ctx, cudaSetDevice(ctx->device));
    const int L = lane_count(B);
    std::vector<nzcb_ctx*> lanes(L);
    for (int l = 0; l < L; l++) {
        lanes[l] = ctx_lane(ctx, l);
        if (!lanes[l]) return ctx->fail(NZCB_E_CUDA, "cannot create lane %d", l);
        lanes[l]->prof_on = ctx->prof_on;
    }
    {   // shared tables exist before any lane needs them
        const Fr* w = nullptr;
        NZ_TRY(get_twiddles_pub(ctx, zk->power, false, &w));
        NZ_TRY(get_twiddles_pub(ctx, zk->power, true, &w));
        NZ_TRY(get_twiddles_pub(ctx, zk->power + 2, false, &w));
        NZ_TRY(get_twiddles_pub(ctx, zk->power + 2, true, &w));
    }
    NZ_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
    for (int l = 0; l < L; l++) NZ_CUDA(ctx, cudaStreamWaitEvent(lanes[l]->stream, ctx->ev0, 0));
    NZ_TRY(prologue(L));  // root-stream work the lanes consume (they wait on its events themselves)
    std::atomic<size_t> next{0};
    std::atomic<int> fatal{0};
    auto worker = [&](int l) {
        nzcb_ctx* lc = lanes[l];
        cudaSetDevice(ctx->device);
        for (;;) {
            const size_t g = next.fetch_add(1);
            if (g >= B || fatal.load()) break;
            const int rc = item(lc, g);
            if (rc == NZCB_E_CUDA || rc == NZCB_E_NOMEM) {
                int expected = 0;
                if (fatal.compare_exchange_strong(expected, rc)) {
                    std::lock_guard<std::mutex> gl(ctx->mu);
                    memcpy(ctx->err, lc->err, sizeof(ctx->err));
                }
                break;
            }
            if (rc != 0) {
                std::lock_guard<std::mutex> gl(ctx->mu);
                memcpy(ctx->err, lc->err, sizeof(ctx->err));
            }
        }
    };
    if (L == 1) {
        worker(0);
    } else {
        std::vector<std::thread> th;
        for (int l = 0; l < L; l++) th.emplace_back(worker, l);
        for (auto& t : th) t.join();
    }
    for (int l = 0; l < L; l++) {
        cudaEventRecord(lanes[l]->ev1, lanes[l]->stream);
        cudaStreamWaitEvent(ctx->stream, lanes[l]->ev1, 0);
    }
    cudaEventRecord(ctx->ev1, ctx->stream);
    const cudaError_t e = cudaStreamSynchronize(ctx->stream);
    for (int l = 0; l < L; l++) cudaStreamSynchronize(lanes[l]->stream);
    if (fatal.load()) return fatal.load();
    if (e != cudaSuccess) return ctx->fail(NZCB_E_CUDA, "CUDA error %s after a batch", cudaGetErrorString(e));
    cudaEventElapsedTime(&ctx->last_ms, ctx->ev0, ctx->ev1);
    return 0;
}

}  // namespace

extern "C" int32_t nzcb_plonk_prove_batch(nzcb_ctx* ctx, const nzcb_zkey* zk, const uint8_t* const* wtns,
                                          const size_t* wtns_len, const uint8_t* blinders_le, size_t B, nzcb_proof* out,
                                          uint8_t* public_le, int32_t* status) {
    if (!ctx || !zk || !wtns || !wtns_len || !out) return NZCB_E_INVALID;
    if (zk->ctx != ctx) return ctx->fail(NZCB_E_INVALID, "zkey was loaded on a different context");
    if (B == 0) return 0;
    std::atomic<int> first_err{0};
    const int rc = run_on_lanes(ctx, zk, B, [&](nzcb_ctx* lc, size_t i) {
        const int r = prove_one(lc, zk, wtns[i], wtns_len[i], blinders_le ? blinders_le + i * 9 * 32 : nullptr, out + i,
                                public_le ? public_le + i * (size_t)zk->n_public * 32 : nullptr);
        if (status) status[i] = r;
        if (r != 0) {
            cudaStreamSynchronize(lc->stream);
            int expected = 0;
            first_err.compare_exchange_strong(expected, r);
            memset(out + i, 0, sizeof(nzcb_proof));
        }
        return r;
    }, [](int) { return 0; });
    if (rc != 0) return rc;
    return status ? 0 : first_err.load();
}

// snarkjs plonk.fullProve: witness program on the GPU, the wires never leave HBM, then the prover
static int32_t fullprove_impl(nzcb_ctx* ctx, const nzcb_circuit* cir, const nzcb_zkey* zk, const uint8_t* inputs_le,
                              bool inputs_on_device, size_t B, const uint8_t* blinders_le, nzcb_proof* out,
                              uint8_t* public_le, int32_t* status);

extern "C" int32_t nzcb_plonk_fullprove_batch(nzcb_ctx* ctx, const nzcb_circuit* cir, const nzcb_zkey* zk,
                                              const uint8_t* inputs_le, size_t B, const uint8_t* blinders_le,
                                              nzcb_proof* out, uint8_t* public_le, int32_t* status) {
    return fullprove_impl(ctx, cir, zk, inputs_le, false, B, blinders_le, out, public_le, status);
}
extern "C" int32_t nzcb_plonk_fullprove_batch_dev(nzcb_ctx* ctx, const nzcb_circuit* cir, const nzcb_zkey* zk,
                                                  const void* d_inputs_le, size_t B, const uint8_t* blinders_le,
                                                  nzcb_proof* out, uint8_t* public_le, int32_t* status) {
    return fullprove_impl(ctx, cir, zk, (const uint8_t*)d_inputs_le, true, B, blinders_le, out, public_le, status);
}

static int32_t fullprove_impl(nzcb_ctx* ctx, const nzcb_circuit* cir, const nzcb_zkey* zk, const uint8_t* inputs_le,
                              bool inputs_on_device, size_t B, const uint8_t* blinders_le, nzcb_proof* out,
                              uint8_t* public_le, int32_t* status) {
    if (!ctx || !cir || !zk || !out || !status || (!inputs_le && circuit_n_in(cir))) return NZCB_E_INVALID;
    if (zk->ctx != ctx) return ctx->fail(NZCB_E_INVALID, "zkey was loaded on a different context");
    const uint32_t n_w = zk->n_vars - zk->n_add;
    if (circuit_n_witness(cir) != n_w)
        return ctx->fail(NZCB_E_WITNESS, "Invalid witness length. Circuit: %u, witness: %u, %u", zk->n_vars,
                         circuit_n_witness(cir), zk->n_add);
    if (B == 0) return 0;
    const size_t n_total = circuit_n_total(cir), n_in = circuit_n_in(cir), n_pub = zk->n_public;
    const size_t per_pass = n_total * sizeof(Fr);
    // The witness program of the whole batch runs on the root stream in two launches: the first `lanes` passes, then
    // the rest -- so the lanes start proving after one witness latency and the second launch hides behind them.
    size_t cap = std::max<size_t>(1, ((size_t)16 << 30) / per_pass);
    int32_t rc_all = 0;
    float total_ms = 0.f;
    for (size_t done = 0; done < B && rc_all == 0; done += cap) {
        const size_t nb = std::min(cap, B - done);
        Fr* d_w = (Fr*)ctx->scratch_get("fp_wires", nb * per_pass);
        Fr* d_in = (Fr*)ctx->scratch_get("fp_inputs", std::max<size_t>(32, nb * n_in * sizeof(Fr)));
        int32_t* d_st = (int32_t*)ctx->scratch_get("fp_status", std::max<size_t>(256, nb * sizeof(int32_t)));
        if (!d_w || !d_in || !d_st) return ctx->fail(NZCB_E_NOMEM, "fullProve: cannot allocate the witness buffers");
        cudaEvent_t ev_first = nullptr, ev_rest = nullptr;
        if (cudaEventCreateWithFlags(&ev_first, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&ev_rest, cudaEventDisableTiming) != cudaSuccess)
            return ctx->fail(NZCB_E_CUDA, "fullProve: cannot create events");
        size_t n_first = 0;
        const uint8_t* in_chunk = inputs_le + done * n_in * 32;
        const uint8_t* bl_chunk = blinders_le ? blinders_le + done * 9 * 32 : nullptr;
        nzcb_proof* out_chunk = out + done;
        uint8_t* pub_chunk = public_le ? public_le + done * n_pub * 32 : nullptr;
        int32_t* st_chunk = status + done;
        rc_all = run_on_lanes(ctx, zk, nb, [&](nzcb_ctx* lc, size_t g) -> int {
            NZ_CUDA(lc, cudaStreamWaitEvent(lc->stream, g < n_first ? ev_first : ev_rest, 0));
            const Fr* w = d_w + g * n_total;
            uint8_t pub[32 * 64];
            std::vector<uint8_t> pub_big;
            uint8_t* pp = pub;
            if (n_pub > 64) {
                pub_big.resize(n_pub * 32);
                pp = pub_big.data();
            }
            int32_t st = 0;
            NZ_CUDA(lc, cudaMemcpyAsync(&st, d_st + g, sizeof(int32_t), cudaMemcpyDeviceToHost, lc->stream));
            if (n_pub) NZ_CUDA(lc, cudaMemcpyAsync(pp, w + 1, n_pub * 32, cudaMemcpyDeviceToHost, lc->stream));
            NZ_CUDA(lc, cudaStreamSynchronize(lc->stream));
            st_chunk[g] = st;
            if (st != 0) {  // "Assert Failed": this pass is rejected, the batch goes on
                memset(out_chunk + g, 0, sizeof(nzcb_proof));
                lc->fail(st, "Assert Failed");
                return 0;
            }
            const int rc = prove_core(lc, zk, w, pp, bl_chunk ? bl_chunk + g * 9 * 32 : nullptr, out_chunk + g,
                                      pub_chunk ? pub_chunk + g * n_pub * 32 : nullptr);
            st_chunk[g] = rc;
            if (rc != 0) {
                cudaStreamSynchronize(lc->stream);
                memset(out_chunk + g, 0, sizeof(nzcb_proof));
            }
            return rc;
        }, [&](int lanes) -> int {
            const Fr* cur_in = d_in;
            if (inputs_on_device) cur_in = (const Fr*)in_chunk;
            else if (n_in)
                NZ_CUDA(ctx, cudaMemcpyAsync(d_in, in_chunk, nb * n_in * 32, cudaMemcpyHostToDevice, ctx->stream));
            n_first = std::min<size_t>(nb, (size_t)lanes);
            NZ_TRY(witness_dev(ctx, cir, cur_in, n_first, d_w, d_st));
            NZ_CUDA(ctx, cudaEventRecord(ev_first, ctx->stream));
            if (nb > n_first)
                NZ_TRY(witness_dev(ctx, cir, cur_in + n_first * n_in, nb - n_first, d_w + n_first * n_total, d_st + n_first));
            NZ_CUDA(ctx, cudaEventRecord(ev_rest, ctx->stream));
            return 0;
        });
        cudaEventDestroy(ev_first);
        cudaEventDestroy(ev_rest);
        total_ms += ctx->last_ms;
    }
    ctx->last_ms = total_ms;
    return rc_all;
}

// ------------------------------------------------------------------ proof.json
namespace {
// 256-bit big-endian bytes -> decimal string
std::string be_to_dec(const uint8_t* be, size_t len) {
    std::vector<uint32_t> limbs((len + 3) / 4, 0);  // little-endian base 2^32
    for (size_t i = 0; i < len; i++) limbs[(len - 1 - i) / 4] |= (uint32_t)be[i] << (8 * ((len - 1 - i) % 4));
    std::string s;
    bool nonzero = true;
    while (nonzero) {
        uint64_t rem = 0;
        nonzero = false;
        for (size_t k = limbs.size(); k-- > 0;) {
            const uint64_t cur = (rem << 32) | limbs[k];
            limbs[k] = (uint32_t)(cur / 1000000000u);
            rem = cur % 1000000000u;
            if (limbs[k]) nonzero = true;
        }
        char buf[16];
        snprintf(buf, sizeof(buf), nonzero ? "%09u" : "%u", (unsigned)rem);
        s = std::string(buf) + s;
    }
    return s;
}
void json_point(std::string& o, const char* key, const uint8_t p[64]) {
    bool inf = true;
    for (int i = 0; i < 64; i++) inf = inf && p[i] == 0;
    o += " \"";
    o += key;
    o += "\": [\n  \"";
    o += inf ? "0" : be_to_dec(p, 32);
    o += "\",\n  \"";
    o += inf ? "1" : be_to_dec(p + 32, 32);
    o += "\",\n  \"";
    o += inf ? "0" : "1";
    o += "\"\n ],\n";
}
void json_scalar(std::string& o, const char* key, const uint8_t p[32]) {
    o += " \"";
    o += key;
    o += "\": \"";
    o += be_to_dec(p, 32);
    o += "\",\n";
}
}  // namespace

extern "C" int32_t nzcb_proof_to_json(const nzcb_proof* p, char* buf, size_t* len) {
    if (!p || !len) return NZCB_E_INVALID;
    std::string o = "{\n";
    json_point(o, "A", p->A);
    json_point(o, "B", p->B);
    json_point(o, "C", p->C);
    json_point(o, "Z", p->Z);
    json_point(o, "T1", p->T1);
    json_point(o, "T2", p->T2);
    json_point(o, "T3", p->T3);
    json_scalar(o, "eval_a", p->eval_a);
    json_scalar(o, "eval_b", p->eval_b);
    json_scalar(o, "eval_c", p->eval_c);
    json_scalar(o, "eval_s1", p->eval_s1);
    json_scalar(o, "eval_s2", p->eval_s2);
    json_scalar(o, "eval_zw", p->eval_zw);
    json_scalar(o, "eval_r", p->eval_r);
    json_point(o, "Wxi", p->Wxi);
    json_point(o, "Wxiw", p->Wxiw);
    o += " \"protocol\": \"plonk\",\n \"curve\": \"bn128\"\n}";
    const size_t need = o.size() + 1;
    if (!buf || *len < need) {
        *len = need;
        return buf ? NZCB_E_INVALID : 0;
    }
    memcpy(buf, o.c_str(), need);
    *len = need;
    return 0;
}
