// Batched affine bucket additions for the MSM (msm.cu step 4a) -- the arithmetic half of what replaces
// ffjavascript G1.multiExpAffine's bucket loop (un-vendored, /root/reference/yarn.lock:3905; nine calls per
// proof).
//
// A bucket accumulated as XYZZ += affine costs 8M + 2S per point.  affine + affine costs one inversion, and
// Montgomery's trick turns the inversions of a batch into one inversion plus three multiplications each:
// 5M + 1S per addition plus a share of the batch's inversion.  The sorted point list is laid out so that every
// bucket's segment starts and ends on a multiple of 2^R entries (the tail of a segment is padded with null
// references); a round then adds entries (2j, 2j+1) into entry j of a list half as long, the segment boundaries
// shift right by one bit, and no round needs to know where the buckets are.  After R rounds the list is
// 2^R times shorter and the XYZZ walk (k_msm_accum) finishes it.
//
// One round = three launches:
//   forward   thread t walks additions [t*M, (t+1)*M): denominators d_j, exclusive prefix products P[j],
//             batch product totals[t]                                                     (1 M per addition)
//   invert    totals[] in place: chunks of 32 per lane, Montgomery's trick + one Fermat inverse per chunk
//   backward  thread t walks its additions in reverse: 1/d_j = inv * P[j]; inv *= d_j; the sum   (2M + 2M + 1S)
// The bodies are host+device so that tests/hostcheck runs the very same code on the CPU.
#pragma once
#include "g1.cuh"

namespace nzcb {

constexpr uint32_t AFF_NULL = 0xffffffffu;  // null point reference: the point at infinity (segment padding)
constexpr uint32_t AFF_M = 32;              // additions per thread and batch
constexpr uint32_t AFF_INV_CHUNK = 32;      // totals per lane in the inversion launch

// entries referenced through the sorted list: bit 31 = negate, low 31 bits = index into the base table
struct AffRefSrc {
    const G1Affine* bases;
    const uint32_t* refs;
    NZ_HD G1Affine get(size_t i) const {
        const uint32_t e = refs[i];
        if (e == AFF_NULL) return G1Affine::inf();
        G1Affine p = bases[e & 0x7fffffffu];
        if (e & 0x80000000u) p.y = p.y.neg();  // (0, 0) stays (0, 0)
        return p;
    }
    // x only (the forward pass needs nothing else unless the two x agree); a null reference reads as x = 0
    NZ_HD Fq get_x(size_t i) const {
        const uint32_t e = refs[i];
        if (e == AFF_NULL) return Fq::zero();
        return bases[e & 0x7fffffffu].x;
    }
};
// entries stored as points (the output of an earlier round)
struct AffPtSrc {
    const G1Affine* pts;
    NZ_HD G1Affine get(size_t i) const { return pts[i]; }
    NZ_HD Fq get_x(size_t i) const { return pts[i].x; }
};

// denominator of the slope of p + q; one() when the sum needs no division (an operand at infinity, p = -q)
NZ_HD Fq aff_denominator(const G1Affine& p, const G1Affine& q) {
    if (p.is_inf() || q.is_inf()) return Fq::one();
    if (p.x == q.x) return p.y == q.y ? p.y.dbl() : Fq::one();  // BN254 G1 has no point with y = 0
    return q.x - p.x;
}

// p + q given dinv = 1 / aff_denominator(p, q)
NZ_HD G1Affine aff_add_with_inv(const G1Affine& p, const G1Affine& q, const Fq& dinv) {
    if (p.is_inf()) return q;
    if (q.is_inf()) return p;
    Fq num;
    if (p.x == q.x) {
        if (p.y != q.y) return G1Affine::inf();
        const Fq xx = p.x.sqr();
        num = xx.dbl() + xx;
    } else {
        num = q.y - p.y;
    }
    const Fq lam = num * dinv;
    G1Affine r;
    r.x = lam.sqr() - p.x - q.x;
    r.y = lam * (p.x - r.x) - p.y;
    return r;
}

template <class Src>
NZ_HD Fq aff_pair_denominator(const Src& src, size_t j) {
    const Fq x1 = src.get_x(2 * j), x2 = src.get_x(2 * j + 1);
    // x = 0 may be the point at infinity, equal x is a doubling or a cancellation: decide on the whole points
    if (x1.is_zero() || x2.is_zero() || x1 == x2) return aff_denominator(src.get(2 * j), src.get(2 * j + 1));
    return x2 - x1;
}

template <class Src>
NZ_HD void aff_forward_body(size_t t, const Src& src, size_t n_add, Fq* P, Fq* totals) {
    const size_t j0 = t * AFF_M;
    if (j0 >= n_add) return;
    const size_t j1 = j0 + AFF_M < n_add ? j0 + AFF_M : n_add;
    Fq prod = aff_pair_denominator(src, j0);
    for (size_t j = j0 + 1; j < j1; j++) {
        P[j] = prod;  // product of the batch's denominators before j (P[j0] is implied: one)
        prod = prod * aff_pair_denominator(src, j);
    }
    totals[t] = prod;
}

template <class Src>
NZ_HD void aff_backward_body(size_t t, const Src& src, size_t n_add, const Fq* P, const Fq* totals_inv, G1Affine* out) {
    const size_t j0 = t * AFF_M;
    if (j0 >= n_add) return;
    const size_t j1 = j0 + AFF_M < n_add ? j0 + AFF_M : n_add;
    Fq inv = totals_inv[t];
    for (size_t j = j1; j-- > j0;) {
        const G1Affine p = src.get(2 * j), q = src.get(2 * j + 1);
        Fq dinv = inv;
        if (j > j0) {
            dinv = inv * P[j];
            inv = inv * aff_denominator(p, q);
        }
        out[j] = aff_add_with_inv(p, q, dinv);
    }
}

// in-place inversion of a[lo .. hi): Montgomery's trick, one Fermat inverse (no element is zero)
NZ_HD void aff_invert_chunk(Fq* a, Fq* tmp, size_t lo, size_t hi) {
    Fq acc = a[lo];
    for (size_t i = lo + 1; i < hi; i++) {
        tmp[i] = acc;
        acc = acc * a[i];
    }
    Fq inv = acc.inv();
    for (size_t i = hi - 1; i > lo; i--) {
        const Fq v = a[i];
        a[i] = inv * tmp[i];
        inv = inv * v;
    }
    a[lo] = inv;
}

}  // namespace nzcb
