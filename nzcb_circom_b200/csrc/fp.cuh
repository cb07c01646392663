// 256-bit prime-field arithmetic for BN254 Fr / Fq, 8 x u32 limbs, Montgomery
// form (R = 2^256).  Replaces wasmcurves' f1m_* / frm_* (un-vendored dependency
// of ffjavascript 0.2.48, /root/reference/yarn.lock:3905,8173; constants per
// SURVEY.md A.1).  Host+device so the very same code is unit-tested on the CPU;
// on the device the multiply is the IMAD carry-chain path below.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define NZ_HD __host__ __device__ __forceinline__
#define NZ_D __device__ __forceinline__
#else
#define NZ_HD inline
#define NZ_D inline
#endif

namespace nzcb {

struct FrParams {
    static NZ_HD uint32_t mod(int i) {
        constexpr uint32_t m[8] = {0xf0000001u, 0x43e1f593u, 0x79b97091u, 0x2833e848u,
                                   0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
        return m[i];
    }
    static NZ_HD uint32_t one(int i) {  // R mod r
        constexpr uint32_t m[8] = {0x4ffffffbu, 0xac96341cu, 0x9f60cd29u, 0x36fc7695u,
                                   0x7879462eu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u};
        return m[i];
    }
    static NZ_HD uint32_t r2(int i) {  // R^2 mod r
        constexpr uint32_t m[8] = {0xae216da7u, 0x1bb8e645u, 0xe35c59e3u, 0x53fe3ab1u,
                                   0x53bb8085u, 0x8c49833du, 0x7f4e44a5u, 0x0216d0b1u};
        return m[i];
    }
    static constexpr uint32_t INV = 0xefffffffu;  // -r^-1 mod 2^32
};

struct FqParams {
    static NZ_HD uint32_t mod(int i) {
        constexpr uint32_t m[8] = {0xd87cfd47u, 0x3c208c16u, 0x6871ca8du, 0x97816a91u,
                                   0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
        return m[i];
    }
    static NZ_HD uint32_t one(int i) {  // R mod p
        constexpr uint32_t m[8] = {0xc58f0d9du, 0xd35d438du, 0xf5c70b3du, 0x0a78eb28u,
                                   0x7879462cu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u};
        return m[i];
    }
    static NZ_HD uint32_t r2(int i) {  // R^2 mod p
        constexpr uint32_t m[8] = {0x538afa89u, 0xf32cfc5bu, 0xd44501fbu, 0xb5e71911u,
                                   0x0a417ff6u, 0x47ab1effu, 0xcab8351fu, 0x06d89f71u};
        return m[i];
    }
    static constexpr uint32_t INV = 0xe4866389u;  // -p^-1 mod 2^32
};


#if defined(__CUDA_ARCH__) && !defined(NZ_NO_PTX_MUL)
#define NZ_PTX_MUL 1
// ---- IMAD carry-chain Montgomery multiply (device) -------------------------
// Two accumulators: E holds words at positions 0..7, O holds positions 1..8
// (O[k] <-> position k+1).  A 32x32 product of an even-indexed limb lands on
// an aligned (E[j], E[j+1]) pair, one of an odd-indexed limb on (O[j-1], O[j]),
// so every row is two unbroken mad.lo.cc / madc.hi.cc chains with no carry
// fix-ups in between.  After a reduction row E[0] == 0 and the word shift is
// free: O becomes the next row's E, and the old E (one word too low for an
// "O") is folded in while the next row is accumulated.
//
// first row: E/O = a * b0 (no carries needed, disjoint words)
NZ_D void nz_mul_row0(uint32_t* E, uint32_t* O, const uint32_t* a, uint32_t bi) {
#pragma unroll
    for (int j = 0; j < 8; j += 2) {
        E[j] = a[j] * bi;
        E[j + 1] = __umulhi(a[j], bi);
        O[j] = a[j + 1] * bi;
        O[j + 1] = __umulhi(a[j + 1], bi);
    }
}
// later rows.  In: E[0..7] at positions 0..7; O[1..7] at positions 0..6 (O[0] dead).
// Out: E, O hold (old value + a * bi) in the standard E/O layout.
NZ_D void nz_mad_row(uint32_t* E, uint32_t* O, const uint32_t* a, uint32_t bi) {
    asm("add.cc.u32      %0, %0, %9;\n\t"
        "madc.lo.cc.u32  %8, %17, %24, %10;\n\t"
        "madc.hi.cc.u32  %9, %17, %24, %11;\n\t"
        "madc.lo.cc.u32  %10, %19, %24, %12;\n\t"
        "madc.hi.cc.u32  %11, %19, %24, %13;\n\t"
        "madc.lo.cc.u32  %12, %21, %24, %14;\n\t"
        "madc.hi.cc.u32  %13, %21, %24, %15;\n\t"
        "madc.lo.cc.u32  %14, %23, %24, 0;\n\t"
        "madc.hi.u32     %15, %23, %24, 0;\n\t"
        "mad.lo.cc.u32   %0, %16, %24, %0;\n\t"
        "madc.hi.cc.u32  %1, %16, %24, %1;\n\t"
        "madc.lo.cc.u32  %2, %18, %24, %2;\n\t"
        "madc.hi.cc.u32  %3, %18, %24, %3;\n\t"
        "madc.lo.cc.u32  %4, %20, %24, %4;\n\t"
        "madc.hi.cc.u32  %5, %20, %24, %5;\n\t"
        "madc.lo.cc.u32  %6, %22, %24, %6;\n\t"
        "madc.hi.cc.u32  %7, %22, %24, %7;\n\t"
        "addc.u32        %15, %15, 0;"
        : "+r"(E[0]), "+r"(E[1]), "+r"(E[2]), "+r"(E[3]), "+r"(E[4]), "+r"(E[5]), "+r"(E[6]), "+r"(E[7]),
          "+r"(O[0]), "+r"(O[1]), "+r"(O[2]), "+r"(O[3]), "+r"(O[4]), "+r"(O[5]), "+r"(O[6]), "+r"(O[7])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(bi));
}
// reduction row: adds mi * modulus with mi = E[0] * INV; leaves E[0] == 0
template <class P>
NZ_D void nz_redc_row(uint32_t* E, uint32_t* O) {
    const uint32_t mi = E[0] * P::INV;
    asm("mad.lo.cc.u32   %8, %17, %24, %8;\n\t"
        "madc.hi.cc.u32  %9, %17, %24, %9;\n\t"
        "madc.lo.cc.u32  %10, %19, %24, %10;\n\t"
        "madc.hi.cc.u32  %11, %19, %24, %11;\n\t"
        "madc.lo.cc.u32  %12, %21, %24, %12;\n\t"
        "madc.hi.cc.u32  %13, %21, %24, %13;\n\t"
        "madc.lo.cc.u32  %14, %23, %24, %14;\n\t"
        "madc.hi.u32     %15, %23, %24, %15;\n\t"
        "mad.lo.cc.u32   %0, %16, %24, %0;\n\t"
        "madc.hi.cc.u32  %1, %16, %24, %1;\n\t"
        "madc.lo.cc.u32  %2, %18, %24, %2;\n\t"
        "madc.hi.cc.u32  %3, %18, %24, %3;\n\t"
        "madc.lo.cc.u32  %4, %20, %24, %4;\n\t"
        "madc.hi.cc.u32  %5, %20, %24, %5;\n\t"
        "madc.lo.cc.u32  %6, %22, %24, %6;\n\t"
        "madc.hi.cc.u32  %7, %22, %24, %7;\n\t"
        "addc.u32        %15, %15, 0;"
        : "+r"(E[0]), "+r"(E[1]), "+r"(E[2]), "+r"(E[3]), "+r"(E[4]), "+r"(E[5]), "+r"(E[6]), "+r"(E[7]),
          "+r"(O[0]), "+r"(O[1]), "+r"(O[2]), "+r"(O[3]), "+r"(O[4]), "+r"(O[5]), "+r"(O[6]), "+r"(O[7])
        : "r"(P::mod(0)), "r"(P::mod(1)), "r"(P::mod(2)), "r"(P::mod(3)), "r"(P::mod(4)), "r"(P::mod(5)),
          "r"(P::mod(6)), "r"(P::mod(7)), "r"(mi));
}
// r = (E + O>>1 word) : E[k] + O[k+1]
NZ_D void nz_merge(uint32_t* r, const uint32_t* E, const uint32_t* O) {
    asm("add.cc.u32   %0, %8, %16;\n\t"
        "addc.cc.u32  %1, %9, %17;\n\t"
        "addc.cc.u32  %2, %10, %18;\n\t"
        "addc.cc.u32  %3, %11, %19;\n\t"
        "addc.cc.u32  %4, %12, %20;\n\t"
        "addc.cc.u32  %5, %13, %21;\n\t"
        "addc.cc.u32  %6, %14, %22;\n\t"
        "addc.u32     %7, %15, 0;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
        : "r"(E[0]), "r"(E[1]), "r"(E[2]), "r"(E[3]), "r"(E[4]), "r"(E[5]), "r"(E[6]), "r"(E[7]),
          "r"(O[1]), "r"(O[2]), "r"(O[3]), "r"(O[4]), "r"(O[5]), "r"(O[6]), "r"(O[7]));
}
#endif

template <class P>
struct alignas(32) Fp {
    uint32_t v[8];

    static NZ_HD Fp zero() {
        Fp r;
#pragma unroll
        for (int i = 0; i < 8; i++) r.v[i] = 0;
        return r;
    }
    static NZ_HD Fp one() {
        Fp r;
#pragma unroll
        for (int i = 0; i < 8; i++) r.v[i] = P::one(i);
        return r;
    }
    static NZ_HD Fp r2() {
        Fp r;
#pragma unroll
        for (int i = 0; i < 8; i++) r.v[i] = P::r2(i);
        return r;
    }
    static NZ_HD Fp modulus() {
        Fp r;
#pragma unroll
        for (int i = 0; i < 8; i++) r.v[i] = P::mod(i);
        return r;
    }
    NZ_HD bool is_zero() const {
        uint32_t o = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) o |= v[i];
        return o == 0;
    }
    NZ_HD bool operator==(const Fp& b) const {
        uint32_t o = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) o |= v[i] ^ b.v[i];
        return o == 0;
    }
    NZ_HD bool operator!=(const Fp& b) const { return !(*this == b); }

    // r = a - mod if a >= mod (a < 2*mod assumed)
    static NZ_HD Fp reduce_once(const Fp& a) {
        Fp t;
        uint32_t borrow = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            uint64_t d = (uint64_t)a.v[i] - P::mod(i) - borrow;
            t.v[i] = (uint32_t)d;
            borrow = (uint32_t)(d >> 63);
        }
        Fp r;
#pragma unroll
        for (int i = 0; i < 8; i++) r.v[i] = borrow ? a.v[i] : t.v[i];
        return r;
    }

    friend NZ_HD Fp operator+(const Fp& a, const Fp& b) {
        Fp s;
        uint32_t carry = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            uint64_t d = (uint64_t)a.v[i] + b.v[i] + carry;
            s.v[i] = (uint32_t)d;
            carry = (uint32_t)(d >> 32);
        }
        return reduce_once(s);  // both moduli < 2^254: no carry out of limb 7
    }
    friend NZ_HD Fp operator-(const Fp& a, const Fp& b) {
        Fp d;
        uint32_t borrow = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            uint64_t t = (uint64_t)a.v[i] - b.v[i] - borrow;
            d.v[i] = (uint32_t)t;
            borrow = (uint32_t)(t >> 63);
        }
        uint32_t mask = 0u - borrow;
        uint32_t carry = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            uint64_t t = (uint64_t)d.v[i] + (P::mod(i) & mask) + carry;
            d.v[i] = (uint32_t)t;
            carry = (uint32_t)(t >> 32);
        }
        return d;
    }
    NZ_HD Fp neg() const { return is_zero() ? *this : modulus() - *this; }
    NZ_HD Fp dbl() const { return *this + *this; }

    // Montgomery product a*b*R^-1 mod m.  CIOS, 8 rows; each row is 8 multiply-
    // accumulates for the operand plus 8 for the reduction (+1 for the quotient
    // digit) -- the 264-IMAD32 unit of SURVEY.md 8(d).  Written on 64-bit
    // accumulators so nvcc emits IMAD.WIDE.U32 chains; carries stay in registers.
    static NZ_HD Fp mul_portable(const Fp& a, const Fp& b) {
        uint32_t t[10];
#pragma unroll
        for (int i = 0; i < 10; i++) t[i] = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            uint64_t c = 0;
            const uint32_t bi = b.v[i];
#pragma unroll
            for (int j = 0; j < 8; j++) {
                c += (uint64_t)a.v[j] * bi + t[j];
                t[j] = (uint32_t)c;
                c >>= 32;
            }
            c += t[8];
            t[8] = (uint32_t)c;
            t[9] = (uint32_t)(c >> 32);
            const uint32_t m = t[0] * P::INV;
            c = (uint64_t)m * P::mod(0) + t[0];
            c >>= 32;
#pragma unroll
            for (int j = 1; j < 8; j++) {
                c += (uint64_t)m * P::mod(j) + t[j];
                t[j - 1] = (uint32_t)c;
                c >>= 32;
            }
            c += t[8];
            t[7] = (uint32_t)c;
            t[8] = t[9] + (uint32_t)(c >> 32);
        }
        Fp r;
#pragma unroll
        for (int i = 0; i < 8; i++) r.v[i] = t[i];
        return reduce_once(r);
    }
#ifdef NZ_PTX_MUL
    static NZ_D Fp mul_ptx(const Fp& a, const Fp& b) {
        uint32_t X[8], Y[8];
        nz_mul_row0(X, Y, a.v, b.v[0]);
        nz_redc_row<P>(X, Y);
#pragma unroll
        for (int i = 1; i < 8; i += 2) {
            nz_mad_row(Y, X, a.v, b.v[i]);
            nz_redc_row<P>(Y, X);
            if (i + 1 < 8) {
                nz_mad_row(X, Y, a.v, b.v[i + 1]);
                nz_redc_row<P>(X, Y);
            }
        }
        // after the 8th row the roles are (E = X, O = Y) again
        Fp r;
        nz_merge(r.v, X, Y);
        return reduce_once(r);
    }
    // the same product with the eight rows rolled into a 4-trip loop (two rows per trip, the multiplier limbs
    // rotated through registers): ~1/3 of the code of mul_ptx, for instruction-cache-bound callers
    static NZ_D Fp mul_rolled(const Fp& a, const Fp& b) {
        uint32_t X[8], Y[8];
#pragma unroll
        for (int j = 0; j < 8; j++) X[j] = Y[j] = 0;
        uint32_t b0 = b.v[0], b1 = b.v[1], b2 = b.v[2], b3 = b.v[3], b4 = b.v[4], b5 = b.v[5], b6 = b.v[6], b7 = b.v[7];
#pragma unroll 1
        for (int it = 0; it < 4; it++) {
            // (E = X, O = Y) holds the running value; Y[0] is dead on entry
            nz_mad_row(X, Y, a.v, b0);   // note: row adds a*bi to E/O given E at 0..7, O[1..7] at 0..6
            nz_redc_row<P>(X, Y);
            nz_mad_row(Y, X, a.v, b1);
            nz_redc_row<P>(Y, X);
            const uint32_t t0 = b0, t1 = b1;
            b0 = b2; b1 = b3; b2 = b4; b3 = b5; b4 = b6; b5 = b7; b6 = t0; b7 = t1;
        }
        Fp r;
        nz_merge(r.v, X, Y);
        return reduce_once(r);
    }
    friend NZ_HD Fp operator*(const Fp& a, const Fp& b) { return mul_ptx(a, b); }
#else
    static NZ_HD Fp mul_rolled(const Fp& a, const Fp& b) { return mul_portable(a, b); }
    friend NZ_HD Fp operator*(const Fp& a, const Fp& b) { return mul_portable(a, b); }
#endif
    NZ_HD Fp sqr() const { return *this * *this; }

    NZ_HD Fp to_mont() const { return *this * r2(); }
    NZ_HD Fp from_mont() const {
        Fp o = zero();
        o.v[0] = 1;
        return *this * o;
    }

    // a^e, e given as 8 little-endian u32 limbs (plain integer exponent)
    NZ_HD Fp pow_limbs(const uint32_t* e) const {
        Fp acc = one();
        for (int i = 7; i >= 0; i--) {
            for (int b = 31; b >= 0; b--) {
                acc = acc.sqr();
                if ((e[i] >> b) & 1) acc = acc * *this;
            }
        }
        return acc;
    }
    // Fermat inverse (0 -> 0), Montgomery in / Montgomery out.
    NZ_HD Fp inv() const {
        uint32_t e[8];
#pragma unroll
        for (int i = 0; i < 8; i++) e[i] = P::mod(i);
        e[0] -= 2;  // both moduli end in ...01 / ...47: no borrow
        return pow_limbs(e);
    }
    NZ_HD Fp pow_u64(uint64_t k) const {
        Fp acc = one();
        Fp base = *this;
        while (k) {
            if (k & 1) acc = acc * base;
            base = base.sqr();
            k >>= 1;
        }
        return acc;
    }
    // small integer -> Montgomery
    static NZ_HD Fp from_u64(uint64_t k) {
        Fp o = zero();
        o.v[0] = (uint32_t)k;
        o.v[1] = (uint32_t)(k >> 32);
        return o.to_mont();
    }
};

typedef Fp<FrParams> Fr;
typedef Fp<FqParams> Fq;

// any 256-bit value -> [0, r): 2^256 < 6 r, five conditional subtractions.  The reduction has to come BEFORE a
// Montgomery conversion or any multiply: the device's carry-chain multiply drops carries its operands cannot produce
// when both are below the modulus, so an unreduced operand gives a wrong product there (the portable host multiply
// is more forgiving, which is why a host check alone does not catch it).
NZ_HD Fr fr_reduce_256(Fr x) {
    for (int i = 0; i < 5; i++) x = Fr::reduce_once(x);
    return x;
}

}  // namespace nzcb
