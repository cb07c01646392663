// Key generation on the GPU: the insecure local SRS (`snarkjs powersoftau new`,
// /root/reference/Makefile:64-67) and `snarkjs plonk setup`
// (/root/reference/Makefile:54-62; algorithm SURVEY.md A.3, snarkjs 0.4.12
// plonk_setup.js as recalled -- un-vendored, /root/reference/yarn.lock:7279).
// The R1CS -> PLONK gate expansion and the copy-permutation are integer work
// done once on the host; every polynomial (5 selectors, 3 sigmas, the Lagrange
// polynomials: n coefficients + 4n evaluations each) and the 8 key commitments
// are computed on the device with the prover's own NTT / MSM kernels.
#include "common.cuh"
#include "poly.cuh"
#include <algorithm>
#include <deque>

using namespace nzcb;

namespace {

// ---------------------------------------------------------------- SRS
__global__ void __launch_bounds__(128) k_srs(Fr tau, size_t count, G1Affine* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const Fr e = tau.pow_u64(i).from_mont();  // canonical scalar tau^i
    G1Affine g;
    g.x = Fq::one();
    g.y = Fq::one() + Fq::one();  // generator (1, 2)
    G1XYZZ acc = G1XYZZ::inf();
    bool started = false;
    for (int l = 7; l >= 0; l--) {
        for (int bit = 31; bit >= 0; bit--) {
            if (started) acc = acc.dbl();
            if ((e.v[l] >> bit) & 1) {
                acc.add_affine(g);
                started = true;
            }
        }
    }
    out[i] = acc.to_affine();
}

__global__ void k_pad4(const Fr* __restrict__ src, size_t n, Fr* __restrict__ dst) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 4 * n) return;
    dst[i] = i < n ? src[i] : Fr::zero();
}

// sigma value at position p = id of position src[p]:  k_col * w^row
__global__ void k_sigma_vals(const uint32_t* __restrict__ src, uint32_t n, uint32_t power, const Fr* __restrict__ Wn,
                             Fr k1, Fr k2, Fr* __restrict__ out) {
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= (size_t)3 * n) return;
    const uint32_t q = src[p];
    const uint32_t col = q / n, row = q % n;
    Fr v = domain_pow(Wn, power, row);
    if (col == 1) v = v * k1;
    if (col == 2) v = v * k2;
    out[p] = v;
}

__global__ void k_unit(Fr* __restrict__ a, size_t n, size_t idx) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    a[i] = i == idx ? Fr::one() : Fr::zero();
}

// ---------------------------------------------------------------- R1CS -> PLONK (host)
struct Term {
    uint32_t s;
    Fr c;  // Montgomery
};
typedef std::vector<Term> LC;

struct Plan {
    uint64_t key_hash = 0;
    size_t key_len = 0;
    uint32_t n_public = 0, r1cs_vars = 0, plonk_vars = 0, power = 0;
    std::vector<uint32_t> sl, sr, so;
    std::vector<Fr> q[5];
    std::vector<uint32_t> add_a, add_b;
    std::vector<Fr> add_ac, add_bc;
    size_t zkey_size() const {
        const size_t n = (size_t)1 << power;
        const size_t nlag = n_public > 1 ? n_public : 1;
        size_t s = 12;
        s += 12 + 4;                                            // section 1
        s += 12 + (4 + 32 + 4 + 32 + 20 + 64 + 8 * 64 + 128);   // section 2
        s += 12 + add_a.size() * 72;                            // section 3
        s += 3 * (12 + sl.size() * 4);                          // 4-6
        s += 5 * (12 + 5 * n * 32);                             // 7-11
        s += 12 + 15 * n * 32;                                  // 12
        s += 12 + nlag * 5 * n * 32;                            // 13
        s += 12 + (n + 6) * 64;                                 // 14
        return s;
    }
};

static Plan g_plan;  // cache between the size query and the real call

void lc_normalize(LC& lc) {
    // r1csfile reads an LC into an object: keys ascending, a repeated signal keeps its last value
    std::stable_sort(lc.begin(), lc.end(), [](const Term& a, const Term& b) { return a.s < b.s; });
    LC out;
    for (size_t i = 0; i < lc.size(); i++) {
        if (i + 1 < lc.size() && lc[i + 1].s == lc[i].s) continue;
        if (!lc[i].c.is_zero()) out.push_back(lc[i]);
    }
    lc.swap(out);
}

struct Builder {
    Plan& p;
    explicit Builder(Plan& pl) : p(pl) {}
    void gate(uint32_t sl, uint32_t sr, uint32_t so, const Fr& qm, const Fr& ql, const Fr& qr, const Fr& qo, const Fr& qc) {
        p.sl.push_back(sl);
        p.sr.push_back(sr);
        p.so.push_back(so);
        p.q[0].push_back(qm);
        p.q[1].push_back(ql);
        p.q[2].push_back(qr);
        p.q[3].push_back(qo);
        p.q[4].push_back(qc);
    }
    // returns constant k and at most max_c (signal, coef) terms; longer LCs are folded pairwise
    // from the front of a queue into new "addition" signals
    void reduce(const LC& lc, size_t max_c, Fr& k, std::vector<Term>& res) {
        k = Fr::zero();
        std::deque<Term> cs;
        for (const Term& t : lc) {
            if (t.s == 0) k = k + t.c;
            else if (!t.c.is_zero()) cs.push_back(t);
        }
        const Fr one = Fr::one(), zero = Fr::zero();
        while (cs.size() > max_c) {
            const Term c1 = cs.front();
            cs.pop_front();
            const Term c2 = cs.front();
            cs.pop_front();
            const uint32_t so = p.plonk_vars++;
            gate(c1.s, c2.s, so, zero, c1.c.neg(), c2.c.neg(), one, zero);
            p.add_a.push_back(c1.s);
            p.add_b.push_back(c2.s);
            p.add_ac.push_back(c1.c);
            p.add_bc.push_back(c2.c);
            cs.push_back(Term{so, one});
        }
        res.assign(cs.begin(), cs.end());
        while (res.size() < max_c) res.push_back(Term{0, zero});
    }
    void add_sum(const LC& lc) {
        Fr k;
        std::vector<Term> c;
        reduce(lc, 3, k, c);
        gate(c[0].s, c[1].s, c[2].s, Fr::zero(), c[0].c, c[1].c, c[2].c, k);
    }
    void add_mul(const LC& la, const LC& lb, const LC& lc) {
        Fr ka, kb, kc;
        std::vector<Term> a, b, c;
        reduce(la, 1, ka, a);
        reduce(lb, 1, kb, b);
        reduce(lc, 1, kc, c);
        gate(a[0].s, b[0].s, c[0].s, a[0].c * b[0].c, a[0].c * kb, ka * b[0].c, c[0].c.neg(), ka * kb - kc);
    }
    // k * lc1 - lc2
    static LC join(const LC& lc1, const Fr& k, const LC& lc2) {
        LC out;
        size_t i = 0, j = 0;
        while (i < lc1.size() || j < lc2.size()) {
            Term t;
            if (j >= lc2.size() || (i < lc1.size() && lc1[i].s < lc2[j].s)) {
                t.s = lc1[i].s;
                t.c = k * lc1[i].c;
                i++;
            } else if (i >= lc1.size() || lc2[j].s < lc1[i].s) {
                t.s = lc2[j].s;
                t.c = lc2[j].c.neg();
                j++;
            } else {
                t.s = lc1[i].s;
                t.c = k * lc1[i].c - lc2[j].c;
                i++;
                j++;
            }
            if (!t.c.is_zero()) out.push_back(t);
        }
        return out;
    }
    // "0": no terms, "k": constant only, otherwise number of signal terms
    static int lc_type(const LC& lc, Fr& k) {
        k = Fr::zero();
        int n = 0;
        for (const Term& t : lc) {
            if (t.s == 0) k = k + t.c;
            else n++;
        }
        if (n > 0) return n;
        return k.is_zero() ? 0 : -1;
    }
};

int build_plan(nzcb_ctx* ctx, const uint8_t* r1cs, size_t len, Plan& p) {
    p = Plan();
    if (len < 12 || memcmp(r1cs, "r1cs", 4) != 0) return ctx->fail(NZCB_E_INVALID, "r1cs file: bad magic");
    uint32_t nsec;
    memcpy(&nsec, r1cs + 8, 4);
    const uint8_t *hdr = nullptr, *body = nullptr;
    uint64_t hdr_len = 0, body_len = 0;
    size_t pos = 12;
    for (uint32_t i = 0; i < nsec; i++) {
        if (pos + 12 > len) return ctx->fail(NZCB_E_INVALID, "r1cs file: truncated");
        uint32_t id;
        uint64_t sz;
        memcpy(&id, r1cs + pos, 4);
        memcpy(&sz, r1cs + pos + 4, 8);
        pos += 12;
        if (pos + sz > len) return ctx->fail(NZCB_E_INVALID, "r1cs file: section overruns the file");
        if (id == 1) { hdr = r1cs + pos; hdr_len = sz; }
        if (id == 2) { body = r1cs + pos; body_len = sz; }
        pos += sz;
    }
    if (!hdr || !body || hdr_len < 4 + 32 + 28) return ctx->fail(NZCB_E_INVALID, "r1cs file: missing sections");
    uint32_t n8;
    memcpy(&n8, hdr, 4);
    if (n8 != 32) return ctx->fail(NZCB_E_INVALID, "r1cs file: field is not bn128 Fr");
    for (int i = 0; i < 8; i++) {
        uint32_t w;
        memcpy(&w, hdr + 4 + 4 * i, 4);
        if (w != FrParams::mod(i)) return ctx->fail(NZCB_E_INVALID, "r1cs file: field is not bn128 Fr");
    }
    uint32_t n_wires, n_out, n_pub_in, n_prv, n_cons;
    memcpy(&n_wires, hdr + 36, 4);
    memcpy(&n_out, hdr + 40, 4);
    memcpy(&n_pub_in, hdr + 44, 4);
    memcpy(&n_prv, hdr + 48, 4);
    memcpy(&n_cons, hdr + 60, 4);
    p.n_public = n_out + n_pub_in;
    p.r1cs_vars = n_wires;
    p.plonk_vars = n_wires;
    Builder bld(p);
    const Fr one = Fr::one(), zero = Fr::zero();
    for (uint32_t s = 1; s <= p.n_public; s++) bld.gate(s, 0, 0, zero, one, zero, zero, zero);
    size_t bp = 0;
    LC lcs[3];
    for (uint32_t c = 0; c < n_cons; c++) {
        for (int k = 0; k < 3; k++) {
            if (bp + 4 > body_len) return ctx->fail(NZCB_E_INVALID, "r1cs file: truncated constraint %u", c);
            uint32_t nt;
            memcpy(&nt, body + bp, 4);
            bp += 4;
            if (bp + (uint64_t)nt * 36 > body_len) return ctx->fail(NZCB_E_INVALID, "r1cs file: truncated constraint %u", c);
            lcs[k].clear();
            lcs[k].reserve(nt);
            for (uint32_t t = 0; t < nt; t++) {
                Term tm;
                memcpy(&tm.s, body + bp, 4);
                Fr v;
                memcpy(v.v, body + bp + 4, 32);
                tm.c = v.to_mont();
                bp += 36;
                if (tm.s >= n_wires) return ctx->fail(NZCB_E_INVALID, "r1cs file: wire index out of range");
                lcs[k].push_back(tm);
            }
            lc_normalize(lcs[k]);
        }
        Fr ka, kb;
        const int ta = Builder::lc_type(lcs[0], ka), tb = Builder::lc_type(lcs[1], kb);
        if (ta == 0 || tb == 0) bld.add_sum(lcs[2]);
        else if (ta == -1) bld.add_sum(Builder::join(lcs[1], ka, lcs[2]));
        else if (tb == -1) bld.add_sum(Builder::join(lcs[0], kb, lcs[2]));
        else bld.add_mul(lcs[0], lcs[1], lcs[2]);
    }
    const size_t ng = p.sl.size();
    uint32_t power = 0;
    while (((size_t)1 << power) < ng) power++;  // = log2(ng - 1) + 1 for ng >= 2
    if (power < 3) power = 3;
    p.power = power;
    p.key_hash = 0;
    p.key_len = len;
    return 0;
}

// cheap identity of an r1cs buffer (length + FNV-1a over a strided sample) so the size query and
// the real call of nzcb_plonk_setup share one gate expansion even when the caller re-marshals
uint64_t r1cs_fingerprint(const uint8_t* d, size_t len) {
    uint64_t h = 1469598103934665603ull ^ len;
    const size_t step = len > (1u << 20) ? 4099 : 1;
    for (size_t i = 0; i < len; i += step) h = (h ^ d[i]) * 1099511628211ull;
    for (size_t i = len > 4096 ? len - 4096 : 0; i < len; i++) h = (h ^ d[i]) * 1099511628211ull;
    return h;
}

int get_plan(nzcb_ctx* ctx, const uint8_t* r1cs, size_t len, Plan** out) {
    const uint64_t fp = r1cs_fingerprint(r1cs, len);
    if (g_plan.key_len != len || g_plan.key_hash != fp || g_plan.sl.empty()) {
        NZ_TRY(build_plan(ctx, r1cs, len, g_plan));
        g_plan.key_hash = fp;
    }
    *out = &g_plan;
    return 0;
}

struct Writer {
    uint8_t* p;
    size_t pos = 0;
    void u32(uint32_t v) { memcpy(p + pos, &v, 4); pos += 4; }
    void u64(uint64_t v) { memcpy(p + pos, &v, 8); pos += 8; }
    void raw(const void* s, size_t n) { memcpy(p + pos, s, n); pos += n; }
    void sec(uint32_t id, uint64_t size) { u32(id); u64(size); }
};

}  // namespace

extern "C" int32_t nzcb_srs_g1(nzcb_ctx* ctx, const uint8_t tau_le[32], size_t count, uint8_t* out) {
    if (!ctx || !tau_le || (count && !out)) return NZCB_E_INVALID;
    if (count == 0) return 0;
    Fr tau;
    memcpy(tau.v, tau_le, 32);
    tau = tau.to_mont();
    G1Affine* d = (G1Affine*)ctx->scratch_get("srs_out", count * sizeof(G1Affine));
    if (!d) return ctx->fail(NZCB_E_NOMEM, "srs: cannot allocate %zu points", count);
    NZ_LAUNCH(ctx, k_srs, div_up(count, 128), 128, 0, tau, count, d);
    NZ_CUDA(ctx, cudaMemcpyAsync(out, d, count * sizeof(G1Affine), cudaMemcpyDeviceToHost, ctx->stream));
    NZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

extern "C" int32_t nzcb_plonk_setup_info(nzcb_ctx* ctx, const uint8_t* r1cs, size_t r1cs_len, uint32_t* n_gates,
                                         uint32_t* n_additions, uint32_t* plonk_vars, uint32_t* power) {
    if (!ctx || !r1cs) return NZCB_E_INVALID;
    Plan* p = nullptr;
    NZ_TRY(get_plan(ctx, r1cs, r1cs_len, &p));
    if (n_gates) *n_gates = (uint32_t)p->sl.size();
    if (n_additions) *n_additions = (uint32_t)p->add_a.size();
    if (plonk_vars) *plonk_vars = p->plonk_vars;
    if (power) *power = p->power;
    return 0;
}

extern "C" int32_t nzcb_plonk_setup(nzcb_ctx* ctx, const uint8_t* r1cs, size_t r1cs_len, const uint8_t* srs,
                                    size_t srs_count, const uint8_t x2[128], uint8_t* zkey_out, size_t* zkey_len) {
    if (!ctx || !r1cs || !zkey_len) return NZCB_E_INVALID;
    Plan* pp = nullptr;
    NZ_TRY(get_plan(ctx, r1cs, r1cs_len, &pp));
    Plan& p = *pp;
    const size_t need = p.zkey_size();
    if (!zkey_out) {
        *zkey_len = need;
        return 0;
    }
    if (*zkey_len < need) {
        *zkey_len = need;
        return ctx->fail(NZCB_E_INVALID, "plonk setup: output buffer too small (need %zu bytes)", need);
    }
    const uint32_t power = p.power;
    const size_t n = (size_t)1 << power;
    const size_t ng = p.sl.size(), na = p.add_a.size();
    if (!srs || srs_count < n + 6)
        return ctx->fail(NZCB_E_INVALID, "circuit too big for this power of tau ceremony. %zu > %zu", n + 6, srs_count);
    NZ_CUDA(ctx, cudaSetDevice(ctx->device));
    // k1 = 2, k2 = 3: the smallest values outside H and outside H u k1 H (valid for every
    // power this build uses since neither 2, 3 nor 3/2 has 2-power order)
    const Fr k1 = Fr::from_u64(2), k2 = Fr::from_u64(3);

    Fr* d_ev = (Fr*)ctx->scratch_get("su_ev", 3 * n * sizeof(Fr));
    Fr* d_ext = (Fr*)ctx->scratch_get("su_ext", 4 * n * sizeof(Fr));
    G1Affine* d_srs = (G1Affine*)ctx->scratch_get("su_srs", (n + 6) * sizeof(G1Affine));
    uint32_t* d_src = (uint32_t*)ctx->scratch_get("su_sig", 3 * n * sizeof(uint32_t));
    G1XYZZ* d_pt = (G1XYZZ*)ctx->scratch_get("su_pt", sizeof(G1XYZZ));
    if (!d_ev || !d_ext || !d_srs || !d_src || !d_pt) return ctx->fail(NZCB_E_NOMEM, "plonk setup: out of device memory");
    NZ_CUDA(ctx, cudaMemcpyAsync(d_srs, srs, (n + 6) * 64, cudaMemcpyHostToDevice, ctx->stream));

    Writer w{zkey_out};
    w.raw("zkey", 4);
    w.u32(1);
    w.u32(14);
    w.sec(1, 4);
    w.u32(2);
    // section 2 header: commitments are patched in after the polynomials exist
    const size_t hdr_size = 4 + 32 + 4 + 32 + 20 + 64 + 8 * 64 + 128;
    w.sec(2, hdr_size);
    const size_t hdr_pos = w.pos;
    {
        uint8_t qle[32], rle[32];
        for (int i = 0; i < 8; i++) {
            const uint32_t a = FqParams::mod(i), b = FrParams::mod(i);
            memcpy(qle + 4 * i, &a, 4);
            memcpy(rle + 4 * i, &b, 4);
        }
        w.u32(32); w.raw(qle, 32); w.u32(32); w.raw(rle, 32);
        w.u32(p.plonk_vars); w.u32(p.n_public); w.u32((uint32_t)n); w.u32((uint32_t)na); w.u32((uint32_t)ng);
        w.raw(k1.v, 32); w.raw(k2.v, 32);
    }
    const size_t commit_pos = w.pos;
    memset(zkey_out + w.pos, 0, 8 * 64);
    w.pos += 8 * 64;
    if (x2) w.raw(x2, 128);
    else { memset(zkey_out + w.pos, 0, 128); w.pos += 128; }
    if (w.pos != hdr_pos + hdr_size) return ctx->fail(NZCB_E_INVALID, "plonk setup: internal header size mismatch");
    // section 3: additions
    w.sec(3, na * 72);
    for (size_t i = 0; i < na; i++) {
        w.u32(p.add_a[i]); w.u32(p.add_b[i]); w.raw(p.add_ac[i].v, 32); w.raw(p.add_bc[i].v, 32);
    }
    w.sec(4, ng * 4); w.raw(p.sl.data(), ng * 4);
    w.sec(5, ng * 4); w.raw(p.sr.data(), ng * 4);
    w.sec(6, ng * 4); w.raw(p.so.data(), ng * 4);

    // one polynomial: evaluations in d_poly (n) -> coefficients + 4n evaluations into the file, optional commitment
    auto emit_poly = [&](Fr* d_poly, int commit_slot) -> int {
        NZ_TRY(ntt_dev(ctx, d_poly, power, true));
        NZ_CUDA(ctx, cudaMemcpyAsync(zkey_out + w.pos, d_poly, n * 32, cudaMemcpyDeviceToHost, ctx->stream));
        w.pos += n * 32;
        NZ_LAUNCH(ctx, k_pad4, div_up(4 * n, 256), 256, 0, d_poly, n, d_ext);
        NZ_TRY(ntt_dev(ctx, d_ext, power + 2, false));
        NZ_CUDA(ctx, cudaMemcpyAsync(zkey_out + w.pos, d_ext, 4 * n * 32, cudaMemcpyDeviceToHost, ctx->stream));
        w.pos += 4 * n * 32;
        if (commit_slot >= 0) {
            NZ_TRY(msm_dev(ctx, d_srs, (const uint32_t*)d_poly, n, true, d_pt));
            G1Affine a;
            NZ_TRY(msm_to_host_affine(ctx, d_pt, &a));
            memcpy(zkey_out + commit_pos + (size_t)commit_slot * 64, &a, 64);
        }
        return 0;
    };

    // sections 7-11: selectors
    for (int k = 0; k < 5; k++) {
        w.sec(7 + k, 5 * n * 32);
        NZ_CUDA(ctx, cudaMemsetAsync(d_ev, 0, n * sizeof(Fr), ctx->stream));
        NZ_CUDA(ctx, cudaMemcpyAsync(d_ev, p.q[k].data(), ng * 32, cudaMemcpyHostToDevice, ctx->stream));
        NZ_TRY(emit_poly(d_ev, k));
    }
    // section 12: sigma.  src[p] = position whose identity value lands at p
    {
        std::vector<uint32_t> src(3 * n), last(p.plonk_vars, 0xffffffffu), first(p.plonk_vars, 0xffffffffu);
        for (size_t i = 0; i < n; i++) {
            const uint32_t sig[3] = {i < ng ? p.sl[i] : 0, i < ng ? p.sr[i] : 0, i < ng ? p.so[i] : 0};
            for (int c = 0; c < 3; c++) {
                const uint32_t pos = (uint32_t)(c * n + i), s = sig[c];
                if (last[s] == 0xffffffffu) first[s] = pos;
                else src[pos] = last[s];
                last[s] = pos;
            }
        }
        for (uint32_t s = 0; s < p.plonk_vars; s++)
            if (first[s] != 0xffffffffu) src[first[s]] = last[s];
        const Fr* Wn = nullptr;
        NZ_TRY(get_twiddles_pub(ctx, power, false, &Wn));
        NZ_CUDA(ctx, cudaMemcpyAsync(d_src, src.data(), 3 * n * 4, cudaMemcpyHostToDevice, ctx->stream));
        NZ_LAUNCH(ctx, k_sigma_vals, div_up(3 * n, 256), 256, 0, d_src, (uint32_t)n, power, Wn, k1, k2, d_ev);
        NZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // src goes out of scope
        w.sec(12, 15 * n * 32);
        for (int c = 0; c < 3; c++) NZ_TRY(emit_poly(d_ev + (size_t)c * n, 5 + c));
    }
    // section 13: Lagrange polynomials L_1 .. L_max(nPublic,1)
    {
        const size_t nlag = p.n_public > 1 ? p.n_public : 1;
        w.sec(13, nlag * 5 * n * 32);
        for (size_t i = 0; i < nlag; i++) {
            NZ_LAUNCH(ctx, k_unit, div_up(n, 256), 256, 0, d_ev, n, i);
            NZ_TRY(emit_poly(d_ev, -1));
        }
    }
    w.sec(14, (n + 6) * 64);
    w.raw(srs, (n + 6) * 64);
    NZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (w.pos != need) return ctx->fail(NZCB_E_INVALID, "plonk setup: internal size mismatch (%zu vs %zu)", w.pos, need);
    *zkey_len = need;
    g_plan = Plan();
    return 0;
}
