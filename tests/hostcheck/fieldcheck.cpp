// Host-side compile of the product's field / curve headers (fp.cuh, g1.cuh) so
// the exact arithmetic the kernels run is checked against the Python oracle
// without a GPU.  Test infrastructure only.
#include <string.h>
#include "../../nzcb_circom_b200/csrc/g1.cuh"
using namespace nzcb;

template <class F> static F ld(const uint8_t* p) { F f; memcpy(f.v, p, 32); return f; }
template <class F> static void st(uint8_t* p, const F& f) { memcpy(p, f.v, 32); }

extern "C" {
// op: 0 add 1 sub 2 mul 3 inv 4 to_mont 5 from_mont 6 neg ; field: 0 Fr 1 Fq
int fc_field_op(int field, int op, const uint8_t* a, const uint8_t* b, uint8_t* out) {
    if (field == 0) {
        Fr x = ld<Fr>(a), y = ld<Fr>(b), r;
        switch (op) { case 0: r = x + y; break; case 1: r = x - y; break; case 2: r = x * y; break;
            case 3: r = x.inv(); break; case 4: r = x.to_mont(); break; case 5: r = x.from_mont(); break;
            case 6: r = x.neg(); break; default: return -1; }
        st(out, r);
    } else {
        Fq x = ld<Fq>(a), y = ld<Fq>(b), r;
        switch (op) { case 0: r = x + y; break; case 1: r = x - y; break; case 2: r = x * y; break;
            case 3: r = x.inv(); break; case 4: r = x.to_mont(); break; case 5: r = x.from_mont(); break;
            case 6: r = x.neg(); break; default: return -1; }
        st(out, r);
    }
    return 0;
}
// points: affine LEM 64 B.  op: 0 = a+b via XYZZ mixed add, 1 = a+b via full add, 2 = dbl(a), 3 = k*a (k u64)
int fc_g1_op(int op, const uint8_t* a, const uint8_t* b, uint64_t k, uint8_t* out) {
    G1Affine A, B; memcpy(&A, a, 64); memcpy(&B, b, 64);
    G1XYZZ r;
    switch (op) {
        case 0: r = G1XYZZ::from_affine(A); r.add_affine(B); break;
        case 1: { r = G1XYZZ::from_affine(A).dbl(); G1XYZZ t = G1XYZZ::from_affine(A); t.add_affine(A.neg());
                  r.add(G1XYZZ::from_affine(A).neg()); /* r = A with non-trivial ZZ */ (void)t;
                  G1XYZZ bb = G1XYZZ::from_affine(B).dbl(); bb.add(G1XYZZ::from_affine(B).neg()); r.add(bb); break; }
        case 2: r = G1XYZZ::from_affine(A).dbl(); break;
        case 3: r = g1_mul_small(G1XYZZ::from_affine(A), k); break;
        default: return -1;
    }
    G1Affine o = r.to_affine(); memcpy(out, &o, 64);
    return 0;
}
}

#include "../../nzcb_circom_b200/csrc/keccak.h"
extern "C" void fc_keccak256(const uint8_t* data, size_t len, uint8_t* out) { nzcb::keccak256(data, len, out); }

// ---- pass ingest (csrc/ingest.cuh): the kernel's per-pass steps, run serially ----
#include "../../nzcb_circom_b200/csrc/ingest.cuh"
// uri: the whole pass URI; inputs_out: (8 maxLen + 161) u32 values; returns the per-pass status (0 / -1)
extern "C" int fc_ingest(const uint8_t* uri, uint32_t total, const uint8_t* data20, uint32_t max_len, uint8_t* tbs_out,
                         uint32_t* tbs_len, uint32_t* inputs_out) {
    static uint8_t raw[ING_MAX_RAW + 3];
    static uint8_t tbs[ING_MAX_TBS];
    const uint32_t n = total > 8 ? total - 8 : 0;
    const uint8_t* sym = uri + 8;
    bool bad = n > ING_MAX_CHARS;
    if (!bad)
        for (uint32_t i = 0; i < n; i++) bad = bad || b32_val(sym[i]) < 0;
    const uint32_t n_raw = bad ? 0 : (n * 5 + 7) / 8;
    for (uint32_t j = 0; j < n_raw; j++) raw[j] = b32_out_byte(sym, n, j);
    CoseFields g = {};
    if (!bad) g = parse_cose(raw, n_raw);
    for (uint32_t t = 0; t < max_len; t++) tbs_out[t] = tbs[t] = tbs_byte_at(g, raw, t);
    *tbs_len = g.ok ? g.tbs_len : 0;
    for (uint32_t i = 0; i < 8 * max_len + 161; i++) inputs_out[i] = ingest_input_value(g, tbs, data20, max_len, i);
    return g.ok ? 0 : -1;
}

// ---- pairing (csrc/pairing.cuh) ----
#include "../../nzcb_circom_b200/csrc/pairing.cuh"
// e(P, Q): P 64 B affine LEM, Q 128 B (x.c0 x.c1 y.c0 y.c1 LEM); out = 6 x Fq2 LEM (384 B)
extern "C" void fc_pairing(const uint8_t* p, const uint8_t* q, uint8_t* out, int do_final) {
    G1Affine P; memcpy(&P, p, 64);
    G2Affine Q; memcpy(&Q, q, 128);
    Fq12 f = miller_loop(P, Q);
    if (do_final) f = final_exp(f);
    memcpy(out, &f, 384);
}
extern "C" int fc_g2_mul(const uint8_t* q, const uint8_t* k_le, uint8_t* out) {
    G2Affine Q; memcpy(&Q, q, 128);
    uint32_t k[8]; memcpy(k, k_le, 32);
    G2Affine r = g2_mul_limbs(Q, k);
    memcpy(out, &r, 128);
    return g2_on_curve(r) ? 1 : 0;
}

// ---- plonk.verify (csrc/verify.cuh) ----
#include "../../nzcb_circom_b200/csrc/verify.cuh"
extern "C" void fc_keccak256_hd(const uint8_t* data, uint32_t len, uint8_t* out) {
    KeccakHD k; k.init(); k.update(data, len); k.finish(out);
}
// vk: n_public u32, power u32, k1 k2 w (LEM), 8 x G1 (LEM 64), X2 (LEM 128)
extern "C" int fc_plonk_verify(const uint8_t* vkb, const uint8_t* proof, const uint8_t* pubs, uint32_t n_pub) {
    VkDev vk;
    memcpy(&vk.n_public, vkb, 4); memcpy(&vk.power, vkb + 4, 4);
    memcpy(&vk.k1, vkb + 8, 32); memcpy(&vk.k2, vkb + 40, 32); memcpy(&vk.w, vkb + 72, 32);
    memcpy(vk.Q, vkb + 104, 512); memcpy(&vk.X2, vkb + 616, 128);
    return verify_serial(vk, proof, pubs, n_pub) ? 1 : 0;
}

// ---- batched affine bucket additions (csrc/msm_affine.cuh): the halving rounds' thread bodies, run serially ----
#include "../../nzcb_circom_b200/csrc/msm_affine.cuh"
#include <vector>
// bases: n_bases x 64 B affine LEM.  refs: the padded, bucket-sorted list (n_refs a multiple of 2^R; AFF_NULL padding;
// bit 31 = negate).  Runs R rounds exactly as msm.cu launches them and writes the (n_refs >> R) surviving points.
extern "C" int fc_affine_rounds(const uint8_t* bases, uint32_t n_bases, const uint32_t* refs, uint32_t n_refs, uint32_t R,
                                uint8_t* out) {
    if (n_refs & ((1u << R) - 1)) return -1;
    std::vector<G1Affine> B(n_bases ? n_bases : 1);
    memcpy(B.data(), bases, (size_t)n_bases * 64);
    std::vector<G1Affine> cur, nxt;
    for (uint32_t r = 1; r <= R; r++) {
        const size_t n_add = n_refs >> r;
        const size_t n_thr = (n_add + AFF_M - 1) / AFF_M;
        std::vector<Fq> P(n_add + 1), tot(n_thr + 1), tmp(n_thr + 1);
        nxt.assign(n_add + 1, G1Affine::inf());
        for (size_t t = 0; t < n_thr + 2; t++) {  // two threads beyond the end: they must do nothing
            if (r == 1) aff_forward_body(t, AffRefSrc{B.data(), refs}, n_add, P.data(), tot.data());
            else aff_forward_body(t, AffPtSrc{cur.data()}, n_add, P.data(), tot.data());
        }
        for (size_t s = 0; s * AFF_INV_CHUNK < n_thr; s++) {
            const size_t lo = s * AFF_INV_CHUNK, hi = lo + AFF_INV_CHUNK < n_thr ? lo + AFF_INV_CHUNK : n_thr;
            aff_invert_chunk(tot.data(), tmp.data(), lo, hi);
        }
        for (size_t t = 0; t < n_thr + 2; t++) {
            if (r == 1) aff_backward_body(t, AffRefSrc{B.data(), refs}, n_add, P.data(), tot.data(), nxt.data());
            else aff_backward_body(t, AffPtSrc{cur.data()}, n_add, P.data(), tot.data(), nxt.data());
        }
        cur.swap(nxt);
    }
    if (R == 0) return -1;
    memcpy(out, cur.data(), (size_t)(n_refs >> R) * 64);
    return 0;
}
