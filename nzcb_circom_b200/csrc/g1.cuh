// BN254 G1 (y^2 = x^3 + 3) point arithmetic over Fq in Montgomery form.
// Affine points are the zkey / ptau in-file form (x||y, infinity = all zero,
// SURVEY.md A.4); accumulators are extended-Jacobian "XYZZ" (x = X/ZZ,
// y = Y/ZZZ, ZZ^3 = ZZZ^2) so a bucket += affine point costs 8M + 2S, the
// 10-modmul unit of SURVEY.md 8(d).  Replaces wasmcurves' g1m_* used by
// ffjavascript G1.multiExpAffine (un-vendored, /root/reference/yarn.lock:3905).
#pragma once
#include "fp.cuh"

namespace nzcb {

struct alignas(32) G1Affine {
    Fq x, y;
    NZ_HD bool is_inf() const { return x.is_zero() && y.is_zero(); }
    static NZ_HD G1Affine inf() {
        G1Affine p;
        p.x = Fq::zero();
        p.y = Fq::zero();
        return p;
    }
    NZ_HD G1Affine neg() const {
        G1Affine p;
        p.x = x;
        p.y = y.neg();
        return p;
    }
};

struct alignas(32) G1XYZZ {
    Fq X, Y, ZZ, ZZZ;

    NZ_HD bool is_inf() const { return ZZ.is_zero(); }
    static NZ_HD G1XYZZ inf() {
        G1XYZZ p;
        p.X = Fq::one();
        p.Y = Fq::one();
        p.ZZ = Fq::zero();
        p.ZZZ = Fq::zero();
        return p;
    }
    static NZ_HD G1XYZZ from_affine(const G1Affine& a) {
        if (a.is_inf()) return inf();
        G1XYZZ p;
        p.X = a.x;
        p.Y = a.y;
        p.ZZ = Fq::one();
        p.ZZZ = Fq::one();
        return p;
    }

    // dbl-2008-s-1 (a = 0)
    NZ_HD G1XYZZ dbl() const {
        if (is_inf()) return *this;
        Fq U = Y.dbl();
        Fq V = U.sqr();
        Fq W = U * V;
        Fq S = X * V;
        Fq X2 = X.sqr();
        Fq M = X2.dbl() + X2;
        G1XYZZ r;
        r.X = M.sqr() - S.dbl();
        r.Y = M * (S - r.X) - W * Y;
        r.ZZ = V * ZZ;
        r.ZZZ = W * ZZZ;
        return r;
    }

    // madd-2008-s: this += affine b   (8M + 2S)
    NZ_HD void add_affine(const G1Affine& b) {
        if (b.is_inf()) return;
        if (is_inf()) {
            *this = from_affine(b);
            return;
        }
        Fq U2 = b.x * ZZ;
        Fq S2 = b.y * ZZZ;
        Fq Pp = U2 - X;
        Fq Rr = S2 - Y;
        if (Pp.is_zero()) {
            if (Rr.is_zero()) {
                *this = from_affine(b).dbl();
            } else {
                *this = inf();
            }
            return;
        }
        Fq PP = Pp.sqr();
        Fq PPP = Pp * PP;
        Fq Q = X * PP;
        Fq X3 = Rr.sqr() - PPP - Q.dbl();
        Fq Y3 = Rr * (Q - X3) - Y * PPP;
        X = X3;
        Y = Y3;
        ZZ = ZZ * PP;
        ZZZ = ZZZ * PPP;
    }

    // add-2008-s: this += b   (12M + 2S)
    NZ_HD void add(const G1XYZZ& b) {
        if (b.is_inf()) return;
        if (is_inf()) {
            *this = b;
            return;
        }
        Fq U1 = X * b.ZZ;
        Fq U2 = b.X * ZZ;
        Fq S1 = Y * b.ZZZ;
        Fq S2 = b.Y * ZZZ;
        Fq Pp = U2 - U1;
        Fq Rr = S2 - S1;
        if (Pp.is_zero()) {
            if (Rr.is_zero()) {
                *this = dbl();
            } else {
                *this = inf();
            }
            return;
        }
        Fq PP = Pp.sqr();
        Fq PPP = Pp * PP;
        Fq Q = U1 * PP;
        Fq X3 = Rr.sqr() - PPP - Q.dbl();
        Fq Y3 = Rr * (Q - X3) - S1 * PPP;
        X = X3;
        Y = Y3;
        ZZ = ZZ * b.ZZ * PP;
        ZZZ = ZZZ * b.ZZZ * PPP;
    }

    NZ_HD G1XYZZ neg() const {
        G1XYZZ r = *this;
        r.Y = Y.neg();
        return r;
    }

    NZ_HD G1Affine to_affine() const {
        if (is_inf()) return G1Affine::inf();
        Fq i = (ZZ * ZZZ).inv();
        G1Affine a;
        a.x = X * (i * ZZZ);
        a.y = Y * (i * ZZ);
        return a;
    }
};

// k * P for a small non-negative integer k (double-and-add, MSB first)
NZ_HD G1XYZZ g1_mul_small(const G1XYZZ& p, uint64_t k) {
    G1XYZZ acc = G1XYZZ::inf();
    for (int b = 63; b >= 0; b--) {
        acc = acc.dbl();
        if ((k >> b) & 1) acc.add(p);
    }
    return acc;
}

}  // namespace nzcb
