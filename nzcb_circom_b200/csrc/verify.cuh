// snarkjs 0.4.12 plonk.verify(vk, publicSignals, proof) up to the pairing (src/plonk_verify.js of the
// un-vendored dependency, /root/reference/yarn.lock:7279; reached from the reference through the exported verifier,
// /root/reference/Makefile:56-57,61-62): well-formedness, the Fiat-Shamir challenges, the Lagrange evaluations,
// t, and the scalars of D, F, E, A1, B1.  The proof is accepted iff  e(-A1, X_2) e(B1, [1]_2) == 1  with
//   B1 = xi Wxi + u xi w Wxiw + T1 + xi^n T2 + xi^2n T3 + v2 A + v3 B + v4 C + v5 S1 + v6 S2
//        + v1 (ea eb Qm + ea Ql + eb Qr + ec Qo + Qc) + dz Z - ds3 S3 - e [1]_1            (18 terms)
//   A1 = Wxi + u Wxiw                                                                       (2 terms)
// Host + device: tests/hostcheck runs the same code against oracle/plonk.py verify on the CPU.
#pragma once
#include "keccak_hd.cuh"
#include "pairing.cuh"

namespace nzcb {

constexpr int VERIFY_TERMS = 20;  // 18 of B1, then 2 of A1

struct VkDev {
    uint32_t n_public, power, pad[6];
    Fr k1, k2, w;     // Montgomery
    G1Affine Q[8];    // Qm Ql Qr Qo Qc S1 S2 S3 (zkey header order), Montgomery
    G2Affine X2;      // [tau]_2
};

// 32 big-endian bytes -> limbs; false if the value is >= the modulus
template <class F>
NZ_HD bool be_to_limbs_checked(const uint8_t* be, F& out) {
    for (int i = 0; i < 8; i++) {
        const uint8_t* p = be + 28 - 4 * i;
        out.v[i] = ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3];
    }
    const F m = F::modulus();
    for (int i = 7; i >= 0; i--) {
        if (out.v[i] < m.v[i]) return true;
        if (out.v[i] > m.v[i]) return false;
    }
    return false;
}
NZ_HD void fr_to_be(const Fr& mont, uint8_t* be) {
    const Fr c = mont.from_mont();
    for (int i = 0; i < 8; i++) {
        uint8_t* p = be + 28 - 4 * i;
        p[0] = (uint8_t)(c.v[i] >> 24), p[1] = (uint8_t)(c.v[i] >> 16), p[2] = (uint8_t)(c.v[i] >> 8), p[3] = (uint8_t)c.v[i];
    }
}
// hashToFr: the 256-bit big-endian digest reduced mod r
NZ_HD Fr hash_finish_fr(KeccakHD& k) {
    uint8_t d[32];
    k.finish(d);
    Fr x;
    for (int i = 0; i < 8; i++) {
        const uint8_t* p = d + 28 - 4 * i;
        x.v[i] = ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3];
    }
    return fr_reduce_256(x).to_mont();
}

// G1 point of the proof: x||y big-endian canonical, infinity = zeros; G1.isValid
NZ_HD bool proof_point(const uint8_t* be, G1Affine& out) {
    Fq x, y;
    if (!be_to_limbs_checked(be, x) || !be_to_limbs_checked(be + 32, y)) return false;
    if (x.is_zero() && y.is_zero()) {
        out = G1Affine::inf();
        return true;
    }
    out.x = x.to_mont();
    out.y = y.to_mont();
    return out.y.sqr() == out.x.sqr() * out.x + Fq::from_u64(3);
}

// Fills pts[20] (Montgomery affine) and sc[20] (canonical little-endian limbs).  proof: the 800-byte nzcb_proof;
// pubs: n_pub x 32 B little-endian (reduced mod r here, as Fr.fromObject does).  false = not well constructed.
NZ_HDN bool verify_prepare(const VkDev& vk, const uint8_t* proof, const uint8_t* pubs, uint32_t n_pub, G1Affine* pts,
                           Fr* sc) {
    if (n_pub != vk.n_public) return false;
    G1Affine P[9];  // A B C Z T1 T2 T3 Wxi Wxiw
    for (int i = 0; i < 9; i++)
        if (!proof_point(proof + 64 * i, P[i])) return false;
    Fr ev[7];       // a b c s1 s2 zw r
    for (int i = 0; i < 7; i++) {
        if (!be_to_limbs_checked(proof + 576 + 32 * i, ev[i])) return false;
        ev[i] = ev[i].to_mont();
    }
    uint8_t buf[32];
    KeccakHD h;
    // beta = H(publicSignals, A, B, C); gamma = H(beta)
    h.init();
    for (uint32_t i = 0; i < n_pub; i++) {
        Fr p;
        for (int k = 0; k < 8; k++) {
            const uint8_t* q = pubs + 32 * i + 4 * k;
            p.v[k] = (uint32_t)q[0] | ((uint32_t)q[1] << 8) | ((uint32_t)q[2] << 16) | ((uint32_t)q[3] << 24);
        }
        fr_to_be(fr_reduce_256(p).to_mont(), buf);
        h.update(buf, 32);
    }
    h.update(proof, 192);
    const Fr beta = hash_finish_fr(h);
    h.init();
    fr_to_be(beta, buf);
    h.update(buf, 32);
    const Fr gamma = hash_finish_fr(h);
    h.init();
    h.update(proof + 192, 64);  // Z
    const Fr alpha = hash_finish_fr(h);
    h.init();
    h.update(proof + 256, 192);  // T1 T2 T3
    const Fr xi = hash_finish_fr(h);
    h.init();
    h.update(proof + 576, 224);  // the seven evaluations
    const Fr v1 = hash_finish_fr(h);
    h.init();
    h.update(proof + 448, 128);  // Wxi Wxiw
    const Fr u = hash_finish_fr(h);
    Fr v[7];
    v[1] = v1;
    for (int i = 2; i < 7; i++) v[i] = v[i - 1] * v1;

    Fr xin = xi;
    for (uint32_t i = 0; i < vk.power; i++) xin = xin.sqr();
    const Fr one = Fr::one();
    const Fr zh = xin - one;
    // L_i(xi) = w^i zh / (n (xi - w^i)), i < max(1, nPublic);  pl = -sum pub_i L_i
    Fr nfr = one;
    for (uint32_t i = 0; i < vk.power; i++) nfr = nfr.dbl();
    Fr wi = one, lag0 = Fr::zero(), pl = Fr::zero();
    const uint32_t n_lag = n_pub ? n_pub : 1;
    for (uint32_t i = 0; i < n_lag; i++) {
        const Fr li = wi * zh * (nfr * (xi - wi)).inv();
        if (i == 0) lag0 = li;
        if (i < n_pub) {
            Fr p;
            for (int k = 0; k < 8; k++) {
                const uint8_t* q = pubs + 32 * i + 4 * k;
                p.v[k] = (uint32_t)q[0] | ((uint32_t)q[1] << 8) | ((uint32_t)q[2] << 16) | ((uint32_t)q[3] << 24);
            }
            pl = pl - fr_reduce_256(p).to_mont() * li;
        }
        wi = wi * vk.w;
    }
    const Fr ea = ev[0], eb = ev[1], ec = ev[2], es1 = ev[3], es2 = ev[4], ezw = ev[5], er = ev[6];
    const Fr alpha2 = alpha.sqr();
    const Fr perm = (ea + beta * es1 + gamma) * (eb + beta * es2 + gamma);
    const Fr t = (er + pl - alpha * perm * (ec + gamma) * ezw - alpha2 * lag0) * zh.inv();
    const Fr bxi = beta * xi;
    const Fr dz = v1 * (alpha * (ea + bxi + gamma) * (eb + bxi * vk.k1 + gamma) * (ec + bxi * vk.k2 + gamma) + alpha2 * lag0) + u;
    const Fr ds3 = v1 * alpha * beta * ezw * perm;
    const Fr e = t + v1 * er + v[2] * ea + v[3] * eb + v[4] * ec + v[5] * es1 + v[6] * es2 + u * ezw;

    G1Affine gen;
    gen.x = Fq::one();
    gen.y = Fq::from_u64(2);
    const G1Affine* Pp[VERIFY_TERMS] = {&P[7], &P[8], &P[4], &P[5], &P[6], &P[0], &P[1], &P[2], &vk.Q[5], &vk.Q[6],
                                        &vk.Q[0], &vk.Q[1], &vk.Q[2], &vk.Q[3], &vk.Q[4], &P[3], &vk.Q[7], &gen,
                                        &P[7], &P[8]};
    const Fr S[VERIFY_TERMS] = {xi, u * xi * vk.w, one, xin, xin.sqr(), v[2], v[3], v[4], v[5], v[6],
                                v1 * ea * eb, v1 * ea, v1 * eb, v1 * ec, v1, dz, ds3.neg(), e.neg(),
                                one, u};
    for (int i = 0; i < VERIFY_TERMS; i++) {
        pts[i] = *Pp[i];
        sc[i] = S[i].from_mont();
    }
    return true;
}

// k * P, k = 8 canonical little-endian limbs
NZ_HDN G1XYZZ g1_mul_limbs(const G1Affine& P, const Fr& k) {
    G1XYZZ acc = G1XYZZ::inf();
    if (P.is_inf()) return acc;
    int top = 255;
    while (top >= 0 && !((k.v[top >> 5] >> (top & 31)) & 1)) top--;
    for (int i = top; i >= 0; i--) {
        acc = acc.dbl();
        if ((k.v[i >> 5] >> (i & 31)) & 1) acc.add_affine(P);
    }
    return acc;
}

// sum of n XYZZ points as an affine point (negated on request)
NZ_HDN G1Affine g1_sum_affine(const G1XYZZ* a, int n, bool negate) {
    G1XYZZ s = a[0];
#pragma unroll 1
    for (int i = 1; i < n; i++) s.add(a[i]);
    if (negate) s = s.neg();
    return s.to_affine();
}

// the whole check on one thread (the kernel spreads the 20 scalar multiplications and the two Miller loops over
// the lanes of a warp; the host check and the single-thread reference use this)
NZ_HDN bool verify_serial(const VkDev& vk, const uint8_t* proof, const uint8_t* pubs, uint32_t n_pub) {
    G1Affine pts[VERIFY_TERMS];
    Fr sc[VERIFY_TERMS];
    if (!verify_prepare(vk, proof, pubs, n_pub, pts, sc)) return false;
    if (!g2_on_curve(vk.X2)) return false;
    G1XYZZ acc[VERIFY_TERMS];
#pragma unroll 1
    for (int i = 0; i < VERIFY_TERMS; i++) acc[i] = g1_mul_limbs(pts[i], sc[i]);
    const Fq12 f = f12_mul(miller_loop(g1_sum_affine(acc + 18, 2, true), vk.X2),
                           miller_loop(g1_sum_affine(acc, 18, false), g2_generator()));
    return final_exp(f).is_one();
}

}  // namespace nzcb
