/* nzcb_oracle.c -- plain-C CPU restatement of the hot path, multi-threaded with OpenMP.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): the checker for mid-size byte parity and the
 * `cpu_baseline` / `--impl reference` leg of bench.py.  It ports oracle/plonk.py (snarkjs 0.4.12
 * plonk.prove as recalled), oracle/witness_vm.py (circom_runtime role) and the
 * ffjavascript primitives (Fr.fft/ifft, G1.multiExpAffine) -- all un-vendored npm dependencies of
 * the reference (/root/reference/yarn.lock:7279,3905,2496), PARITY UNPINNED at that level; this
 * file is pinned against the Python oracle by tests/test_c_oracle.py.
 * Arithmetic: 4 x 64-bit limbs, Montgomery R = 2^256, unsigned __int128 products.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef unsigned __int128 u128;
/* field element; file sections are only 4-byte aligned, so never let the compiler assume more */
typedef struct __attribute__((packed, aligned(4))) { uint64_t v[4]; } fe;

typedef struct {
    uint64_t m[4], r1[4], r2[4], inv; /* modulus, R mod m, R^2 mod m, -m^-1 mod 2^64 */
} field;

static const field FR = {
    {0x43e1f593f0000001ULL, 0x2833e84879b97091ULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL},
    {0xac96341c4ffffffbULL, 0x36fc76959f60cd29ULL, 0x666ea36f7879462eULL, 0x0e0a77c19a07df2fULL},
    {0x1bb8e645ae216da7ULL, 0x53fe3ab1e35c59e3ULL, 0x8c49833d53bb8085ULL, 0x0216d0b17f4e44a5ULL},
    0xc2e1f593efffffffULL};
static const field FQ = {
    {0x3c208c16d87cfd47ULL, 0x97816a916871ca8dULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL},
    {0xd35d438dc58f0d9dULL, 0x0a78eb28f5c70b3dULL, 0x666ea36f7879462cULL, 0x0e0a77c19a07df2fULL},
    {0xf32cfc5b538afa89ULL, 0xb5e71911d44501fbULL, 0x47ab1eff0a417ff6ULL, 0x06d89f71cab8351fULL},
    0x87d20782e4866389ULL};

static inline int fe_is_zero(const fe* a) { return (a->v[0] | a->v[1] | a->v[2] | a->v[3]) == 0; }
static inline int fe_eq(const fe* a, const fe* b) {
    return ((a->v[0] ^ b->v[0]) | (a->v[1] ^ b->v[1]) | (a->v[2] ^ b->v[2]) | (a->v[3] ^ b->v[3])) == 0;
}
static inline int geq(const uint64_t* a, const uint64_t* m) {
    for (int i = 3; i >= 0; i--) { if (a[i] > m[i]) return 1; if (a[i] < m[i]) return 0; }
    return 1;
}
static inline void sub_n(uint64_t* r, const uint64_t* a, const uint64_t* b) {
    uint64_t br = 0;
    for (int i = 0; i < 4; i++) { u128 d = (u128)a[i] - b[i] - br; r[i] = (uint64_t)d; br = (uint64_t)(d >> 64) & 1; }
}
static inline void f_add(const field* F, fe* r, const fe* a, const fe* b) {
    uint64_t c = 0, t[4];
    for (int i = 0; i < 4; i++) { u128 s = (u128)a->v[i] + b->v[i] + c; t[i] = (uint64_t)s; c = (uint64_t)(s >> 64); }
    if (geq(t, F->m)) sub_n(r->v, t, F->m); else memcpy(r->v, t, 32);
}
static inline void f_sub(const field* F, fe* r, const fe* a, const fe* b) {
    uint64_t br = 0, t[4];
    for (int i = 0; i < 4; i++) { u128 d = (u128)a->v[i] - b->v[i] - br; t[i] = (uint64_t)d; br = (uint64_t)(d >> 64) & 1; }
    if (br) { uint64_t c = 0; for (int i = 0; i < 4; i++) { u128 s = (u128)t[i] + F->m[i] + c; t[i] = (uint64_t)s; c = (uint64_t)(s >> 64); } }
    memcpy(r->v, t, 32);
}
static inline void f_neg(const field* F, fe* r, const fe* a) {
    if (fe_is_zero(a)) { *r = *a; return; }
    sub_n(r->v, F->m, a->v);
}
static inline void f_mul(const field* F, fe* r, const fe* a, const fe* b) {
    uint64_t t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; i++) {
        u128 c = 0;
        for (int j = 0; j < 4; j++) { c += (u128)a->v[j] * b->v[i] + t[j]; t[j] = (uint64_t)c; c >>= 64; }
        c += t[4]; t[4] = (uint64_t)c; t[5] = (uint64_t)(c >> 64);
        uint64_t m = t[0] * F->inv;
        c = (u128)m * F->m[0] + t[0]; c >>= 64;
        for (int j = 1; j < 4; j++) { c += (u128)m * F->m[j] + t[j]; t[j - 1] = (uint64_t)c; c >>= 64; }
        c += t[4]; t[3] = (uint64_t)c; t[4] = t[5] + (uint64_t)(c >> 64);
    }
    if (geq(t, F->m)) sub_n(r->v, t, F->m); else memcpy(r->v, t, 32);
}
static inline void f_sqr(const field* F, fe* r, const fe* a) { f_mul(F, r, a, a); }
static inline void f_one(const field* F, fe* r) { memcpy(r->v, F->r1, 32); }
static inline void f_zero(fe* r) { memset(r->v, 0, 32); }
static inline void f_to_mont(const field* F, fe* r, const fe* a) { fe r2; memcpy(r2.v, F->r2, 32); f_mul(F, r, a, &r2); }
static inline void f_from_mont(const field* F, fe* r, const fe* a) { fe o = {{1, 0, 0, 0}}; f_mul(F, r, a, &o); }
static void f_pow(const field* F, fe* r, const fe* a, const uint64_t e[4]) {
    fe acc; f_one(F, &acc);
    for (int i = 3; i >= 0; i--) for (int b = 63; b >= 0; b--) { f_sqr(F, &acc, &acc); if ((e[i] >> b) & 1) f_mul(F, &acc, &acc, a); }
    *r = acc;
}
static void f_inv(const field* F, fe* r, const fe* a) { uint64_t e[4]; memcpy(e, F->m, 32); e[0] -= 2; f_pow(F, r, a, e); }
static void f_from_u64(const field* F, fe* r, uint64_t k) { fe t = {{k, 0, 0, 0}}; f_to_mont(F, r, &t); }

#define FRADD(r, a, b) f_add(&FR, r, a, b)
#define FRSUB(r, a, b) f_sub(&FR, r, a, b)
#define FRMUL(r, a, b) f_mul(&FR, r, a, b)

/* ---- roots of unity: w[28] = 5^((r-1)/2^28), w[i] = w[i+1]^2 ---- */
static void fr_root(fe* w, unsigned log_n) {
    uint64_t e[4], s[4];
    memcpy(e, FR.m, 32); e[0] -= 1;
    for (int i = 0; i < 4; i++) s[i] = (e[i] >> 28) | (i < 3 ? e[i + 1] << 36 : 0);
    fe five; f_from_u64(&FR, &five, 5);
    f_pow(&FR, w, &five, s);
    for (unsigned i = 28; i > log_n; i--) f_sqr(&FR, w, w);
}

/* ---- NTT (ffjavascript Fr.fft / ifft semantics) ---- */
static void ntt_core(fe* a, unsigned log_n, int inverse) {
    const size_t n = (size_t)1 << log_n;
    if (log_n == 0) return;
    /* bit reversal */
    for (size_t i = 0; i < n; i++) {
        size_t r = 0;
        for (unsigned b = 0; b < log_n; b++) r |= ((i >> b) & 1) << (log_n - 1 - b);
        if (i < r) { fe t = a[i]; a[i] = a[r]; a[r] = t; }
    }
    fe w; fr_root(&w, log_n);
    if (inverse) f_inv(&FR, &w, &w);
    /* twiddle table w^t, t < n/2 */
    fe* W = (fe*)malloc((n / 2 ? n / 2 : 1) * sizeof(fe));
    f_one(&FR, &W[0]);
    for (size_t t = 1; t < n / 2; t++) FRMUL(&W[t], &W[t - 1], &w);
    for (unsigned s = 0; s < log_n; s++) {
        const size_t m = (size_t)1 << s, step = n >> (s + 1);
#pragma omp parallel for schedule(static)
        for (size_t k = 0; k < n / 2; k++) {
            const size_t blk = k >> s, j = k & (m - 1);
            const size_t i0 = (blk << (s + 1)) + j, i1 = i0 + m;
            fe v; FRMUL(&v, &a[i1], &W[j * step]);
            fe u = a[i0];
            FRADD(&a[i0], &u, &v);
            FRSUB(&a[i1], &u, &v);
        }
    }
    if (inverse) {
        fe ninv; f_from_u64(&FR, &ninv, (uint64_t)n); f_inv(&FR, &ninv, &ninv);
#pragma omp parallel for schedule(static)
        for (size_t i = 0; i < n; i++) FRMUL(&a[i], &a[i], &ninv);
    }
    free(W);
}
void oracle_ntt(uint8_t* data_lem, unsigned log_n, int inverse) { ntt_core((fe*)data_lem, log_n, inverse); }

/* ---- G1, Jacobian ---- */
typedef struct { fe x, y; } g1a;        /* affine, Montgomery; (0,0) = infinity */
typedef struct { fe X, Y, Z; } g1j;
static inline int g1a_is_inf(const g1a* p) { return fe_is_zero(&p->x) && fe_is_zero(&p->y); }
static inline void g1j_inf(g1j* p) { f_one(&FQ, &p->X); f_one(&FQ, &p->Y); f_zero(&p->Z); }
static void g1j_dbl(g1j* r, const g1j* p) {
    if (fe_is_zero(&p->Z)) { *r = *p; return; }
    fe A, B, C, D, E, F, t, X3, Y3, Z3;
    f_sqr(&FQ, &A, &p->X); f_sqr(&FQ, &B, &p->Y); f_sqr(&FQ, &C, &B);
    f_add(&FQ, &t, &p->X, &B); f_sqr(&FQ, &t, &t); f_sub(&FQ, &t, &t, &A); f_sub(&FQ, &t, &t, &C); f_add(&FQ, &D, &t, &t);
    f_add(&FQ, &E, &A, &A); f_add(&FQ, &E, &E, &A);
    f_sqr(&FQ, &F, &E);
    f_sub(&FQ, &X3, &F, &D); f_sub(&FQ, &X3, &X3, &D);
    f_sub(&FQ, &t, &D, &X3); f_mul(&FQ, &Y3, &E, &t);
    f_add(&FQ, &t, &C, &C); f_add(&FQ, &t, &t, &t); f_add(&FQ, &t, &t, &t); f_sub(&FQ, &Y3, &Y3, &t);
    f_mul(&FQ, &Z3, &p->Y, &p->Z); f_add(&FQ, &Z3, &Z3, &Z3);
    r->X = X3; r->Y = Y3; r->Z = Z3;
}
static void g1j_add_affine(g1j* r, const g1j* p, const g1a* q) {
    if (g1a_is_inf(q)) { *r = *p; return; }
    if (fe_is_zero(&p->Z)) { r->X = q->x; r->Y = q->y; f_one(&FQ, &r->Z); return; }
    fe Z1Z1, U2, S2, H, HH, I, J, rr, V, t, X3, Y3, Z3;
    f_sqr(&FQ, &Z1Z1, &p->Z); f_mul(&FQ, &U2, &q->x, &Z1Z1);
    f_mul(&FQ, &S2, &q->y, &p->Z); f_mul(&FQ, &S2, &S2, &Z1Z1);
    f_sub(&FQ, &H, &U2, &p->X); f_sub(&FQ, &rr, &S2, &p->Y);
    if (fe_is_zero(&H)) {
        if (fe_is_zero(&rr)) { g1j_dbl(r, p); return; }
        g1j_inf(r); return;
    }
    f_sqr(&FQ, &HH, &H); f_add(&FQ, &I, &HH, &HH); f_add(&FQ, &I, &I, &I);
    f_mul(&FQ, &J, &H, &I); f_add(&FQ, &rr, &rr, &rr); f_mul(&FQ, &V, &p->X, &I);
    f_sqr(&FQ, &X3, &rr); f_sub(&FQ, &X3, &X3, &J); f_sub(&FQ, &X3, &X3, &V); f_sub(&FQ, &X3, &X3, &V);
    f_sub(&FQ, &t, &V, &X3); f_mul(&FQ, &Y3, &rr, &t); f_mul(&FQ, &t, &p->Y, &J); f_add(&FQ, &t, &t, &t); f_sub(&FQ, &Y3, &Y3, &t);
    f_add(&FQ, &Z3, &p->Z, &H); f_sqr(&FQ, &Z3, &Z3); f_sub(&FQ, &Z3, &Z3, &Z1Z1); f_sub(&FQ, &Z3, &Z3, &HH);
    r->X = X3; r->Y = Y3; r->Z = Z3;
}
static void g1j_add(g1j* r, const g1j* p, const g1j* q) {
    if (fe_is_zero(&q->Z)) { *r = *p; return; }
    if (fe_is_zero(&p->Z)) { *r = *q; return; }
    fe Z1Z1, Z2Z2, U1, U2, S1, S2, H, I, J, rr, V, t, X3, Y3, Z3;
    f_sqr(&FQ, &Z1Z1, &p->Z); f_sqr(&FQ, &Z2Z2, &q->Z);
    f_mul(&FQ, &U1, &p->X, &Z2Z2); f_mul(&FQ, &U2, &q->X, &Z1Z1);
    f_mul(&FQ, &S1, &p->Y, &q->Z); f_mul(&FQ, &S1, &S1, &Z2Z2);
    f_mul(&FQ, &S2, &q->Y, &p->Z); f_mul(&FQ, &S2, &S2, &Z1Z1);
    f_sub(&FQ, &H, &U2, &U1); f_sub(&FQ, &rr, &S2, &S1);
    if (fe_is_zero(&H)) {
        if (fe_is_zero(&rr)) { g1j_dbl(r, p); return; }
        g1j_inf(r); return;
    }
    f_add(&FQ, &I, &H, &H); f_sqr(&FQ, &I, &I); f_mul(&FQ, &J, &H, &I); f_add(&FQ, &rr, &rr, &rr); f_mul(&FQ, &V, &U1, &I);
    f_sqr(&FQ, &X3, &rr); f_sub(&FQ, &X3, &X3, &J); f_sub(&FQ, &X3, &X3, &V); f_sub(&FQ, &X3, &X3, &V);
    f_sub(&FQ, &t, &V, &X3); f_mul(&FQ, &Y3, &rr, &t); f_mul(&FQ, &t, &S1, &J); f_add(&FQ, &t, &t, &t); f_sub(&FQ, &Y3, &Y3, &t);
    f_add(&FQ, &Z3, &p->Z, &q->Z); f_sqr(&FQ, &Z3, &Z3); f_sub(&FQ, &Z3, &Z3, &Z1Z1); f_sub(&FQ, &Z3, &Z3, &Z2Z2); f_mul(&FQ, &Z3, &Z3, &H);
    r->X = X3; r->Y = Y3; r->Z = Z3;
}
static void g1j_to_affine(g1a* r, const g1j* p) {
    if (fe_is_zero(&p->Z)) { f_zero(&r->x); f_zero(&r->y); return; }
    fe zi, zi2, zi3; f_inv(&FQ, &zi, &p->Z); f_sqr(&FQ, &zi2, &zi); f_mul(&FQ, &zi3, &zi2, &zi);
    f_mul(&FQ, &r->x, &p->X, &zi2); f_mul(&FQ, &r->y, &p->Y, &zi3);
}

/* MSM: Pippenger, unsigned c-bit windows, one OpenMP task per (window, slice); scalars Montgomery or canonical */
static void msm_core(g1a* out, const g1a* bases, const fe* scalars, size_t n, int mont) {
    g1j total; g1j_inf(&total);
    if (n == 0) { g1j_to_affine(out, &total); return; }
    unsigned lg = 0; while (((size_t)2 << lg) <= n) lg++;
    int c = (int)lg - 3; if (c < 3) c = 3; if (c > 16) c = 16;
    const int nwin = (254 + c - 1) / c;
    fe* sc = (fe*)malloc(n * sizeof(fe));
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) { if (mont) f_from_mont(&FR, &sc[i], &scalars[i]); else sc[i] = scalars[i]; }
    int nthreads = 1;
#ifdef _OPENMP
    nthreads = omp_get_max_threads();
#endif
    int slices = (nthreads + nwin - 1) / nwin; if (slices < 1) slices = 1;
    if ((size_t)slices > n) slices = (int)n;
    g1j* part = (g1j*)malloc((size_t)nwin * slices * sizeof(g1j));
#pragma omp parallel for schedule(dynamic, 1) collapse(2)
    for (int w = 0; w < nwin; w++) {
        for (int s = 0; s < slices; s++) {
            const size_t lo = n * (size_t)s / slices, hi = n * (size_t)(s + 1) / slices;
            const size_t nb = (size_t)1 << c;
            g1j* bk = (g1j*)malloc(nb * sizeof(g1j));
            for (size_t k = 0; k < nb; k++) g1j_inf(&bk[k]);
            const unsigned off = (unsigned)(w * c);
            for (size_t i = lo; i < hi; i++) {
                const unsigned limb = off >> 6, sh = off & 63;
                uint64_t d = sc[i].v[limb] >> sh;
                if (sh + c > 64 && limb + 1 < 4) d |= sc[i].v[limb + 1] << (64 - sh);
                d &= (nb - 1);
                if (d && !g1a_is_inf(&bases[i])) g1j_add_affine(&bk[d], &bk[d], &bases[i]);
            }
            g1j run, acc; g1j_inf(&run); g1j_inf(&acc);
            for (size_t k = nb - 1; k >= 1; k--) { g1j_add(&run, &run, &bk[k]); g1j_add(&acc, &acc, &run); }
            part[(size_t)w * slices + s] = acc;
            free(bk);
        }
    }
    for (int w = nwin - 1; w >= 0; w--) {
        for (int i = 0; i < c; i++) g1j_dbl(&total, &total);
        for (int s = 0; s < slices; s++) g1j_add(&total, &total, &part[(size_t)w * slices + s]);
    }
    g1j_to_affine(out, &total);
    free(part); free(sc);
}
void oracle_msm(const uint8_t* bases_lem, const uint8_t* scalars_le, size_t n, uint8_t* out_lem) {
    msm_core((g1a*)out_lem, (const g1a*)bases_lem, (const fe*)scalars_le, n, 0);
}

/* ---- Keccak-256 + hashToFr ---- */
static uint64_t rol64(uint64_t x, int n) { return n ? (x << n) | (x >> (64 - n)) : x; }
static void keccak_f(uint64_t s[25]) {
    static const int ROT[25] = {0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39, 41, 45, 15, 21, 8, 18, 2, 61, 56, 14};
    uint64_t rc = 1;
    for (int r = 0; r < 24; r++) {
        uint64_t C[5], D[5], B[25];
        for (int x = 0; x < 5; x++) C[x] = s[x] ^ s[x + 5] ^ s[x + 10] ^ s[x + 15] ^ s[x + 20];
        for (int x = 0; x < 5; x++) D[x] = C[(x + 4) % 5] ^ rol64(C[(x + 1) % 5], 1);
        for (int i = 0; i < 25; i++) s[i] ^= D[i % 5];
        for (int x = 0; x < 5; x++) for (int y = 0; y < 5; y++) B[y + 5 * ((2 * x + 3 * y) % 5)] = rol64(s[x + 5 * y], ROT[x + 5 * y]);
        for (int x = 0; x < 5; x++) for (int y = 0; y < 5; y++) s[x + 5 * y] = B[x + 5 * y] ^ ((~B[(x + 1) % 5 + 5 * y]) & B[(x + 2) % 5 + 5 * y]);
        /* round constant from the degree-8 LFSR of the Keccak spec */
        uint64_t c = 0;
        for (int j = 0; j < 7; j++) {
            if (rc & 1) c ^= (uint64_t)1 << ((1 << j) - 1);
            rc = ((rc << 1) ^ ((rc >> 7) * 0x71)) & 0xff;
        }
        s[0] ^= c;
    }
}
void oracle_keccak256(const uint8_t* data, size_t len, uint8_t out[32]) {
    uint64_t s[25]; memset(s, 0, sizeof(s));
    const size_t rate = 136;
    size_t padded = (len / rate + 1) * rate;
    uint8_t* m = (uint8_t*)calloc(padded, 1);
    memcpy(m, data, len); m[len] = 0x01; m[padded - 1] |= 0x80;
    for (size_t off = 0; off < padded; off += rate) {
        for (size_t i = 0; i < rate / 8; i++) { uint64_t v = 0; for (int k = 7; k >= 0; k--) v = (v << 8) | m[off + 8 * i + k]; s[i] ^= v; }
        keccak_f(s);
    }
    for (int i = 0; i < 4; i++) for (int k = 0; k < 8; k++) out[8 * i + k] = (uint8_t)(s[i] >> (8 * k));
    free(m);
}
static void hash_to_fr(fe* r, const uint8_t* msg, size_t len) {
    uint8_t dg[32]; oracle_keccak256(msg, len, dg);
    fe v;
    for (int i = 0; i < 4; i++) { uint64_t w = 0; for (int k = 0; k < 8; k++) w |= (uint64_t)dg[31 - (8 * i + k)] << (8 * k); v.v[i] = w; }
    while (geq(v.v, FR.m)) sub_n(v.v, v.v, FR.m);
    f_to_mont(&FR, r, &v);
}
static void fr_to_be(uint8_t out[32], const fe* mont) {
    fe c; f_from_mont(&FR, &c, mont);
    for (int i = 0; i < 4; i++) for (int k = 0; k < 8; k++) out[31 - (8 * i + k)] = (uint8_t)(c.v[i] >> (8 * k));
}
static void g1_to_be(uint8_t out[64], const g1a* p) {
    fe c;
    f_from_mont(&FQ, &c, &p->x); for (int i = 0; i < 4; i++) for (int k = 0; k < 8; k++) out[31 - (8 * i + k)] = (uint8_t)(c.v[i] >> (8 * k));
    f_from_mont(&FQ, &c, &p->y); for (int i = 0; i < 4; i++) for (int k = 0; k < 8; k++) out[63 - (8 * i + k)] = (uint8_t)(c.v[i] >> (8 * k));
}

/* ---- witness program interpreter (oracle/witness_vm.py) ---- */
enum { OP_LIN = 1, OP_MUL = 2, OP_BITS = 3, OP_INV = 4, OP_ASSERT = 5, OP_BITSLC = 6 };
static fe eval_lc(const uint32_t* code, uint32_t* p, const fe* consts_m, const fe* W) {
    const uint32_t n = code[*p], ci = code[*p + 1];
    *p += 2;
    fe acc; f_zero(&acc);
    if (ci != 0xffffffffu) f_from_mont(&FR, &acc, &consts_m[ci]);
    for (uint32_t t = 0; t < n; t++) {
        const uint32_t w = code[*p], c = code[*p + 1];
        *p += 2;
        if (c == 0) FRADD(&acc, &acc, &W[w]);
        else if (c == 1) FRSUB(&acc, &acc, &W[w]);
        else if (!fe_is_zero(&W[w])) { fe t2; FRMUL(&t2, &consts_m[c], &W[w]); FRADD(&acc, &acc, &t2); }
    }
    return acc;
}
/* returns 0, or -6 (Assert Failed).  wires: n_total canonical LE values (caller allocates) */
int oracle_witness(const uint8_t* wprog, size_t len, const uint8_t* inputs_le, uint8_t* wires_out) {
    if (len < 40 || memcmp(wprog, "NZWP", 4) != 0) return -1;
    uint32_t h[9]; memcpy(h, wprog + 4, 36);
    const uint32_t n_total = h[1], n_out = h[3], n_in = h[4], n_consts = h[5], n_instr = h[6], n_levels = h[7];
    const fe* consts = (const fe*)(wprog + 40);
    const uint32_t* ioff = (const uint32_t*)(wprog + 40 + (size_t)n_consts * 32);
    const uint32_t* code = ioff + n_instr + n_levels + 1;
    fe* cm = (fe*)malloc((size_t)n_consts * sizeof(fe));
    for (uint32_t i = 0; i < n_consts; i++) f_to_mont(&FR, &cm[i], &consts[i]);
    fe* W = (fe*)wires_out;
    memset(W, 0, (size_t)n_total * 32);
    W[0].v[0] = 1;
    memcpy(&W[1 + n_out], inputs_le, (size_t)n_in * 32);
    fe r2; memcpy(r2.v, FR.r2, 32);
    int failed = 0;
    for (uint32_t i = 0; i < n_instr; i++) {
        uint32_t p = ioff[i];
        const uint32_t op = code[p];
        if (op == OP_LIN) { const uint32_t d = code[p + 1]; p += 2; W[d] = eval_lc(code, &p, cm, W); }
        else if (op == OP_MUL || op == OP_ASSERT) {
            uint32_t d = 0;
            if (op == OP_MUL) { d = code[p + 1]; p += 2; } else p += 1;
            fe a = eval_lc(code, &p, cm, W), b = eval_lc(code, &p, cm, W), c = eval_lc(code, &p, cm, W), ab;
            FRMUL(&ab, &a, &b); FRMUL(&ab, &ab, &r2);
            if (op == OP_MUL) FRADD(&W[d], &ab, &c); else if (!fe_eq(&ab, &c)) failed = 1;
        } else if (op == OP_BITS) {
            const uint32_t d = code[p + 1], src = code[p + 2], n = code[p + 3];
            const fe v = W[src];
            for (uint32_t k = 0; k < n; k++) { f_zero(&W[d + k]); W[d + k].v[0] = k < 256 ? (v.v[k >> 6] >> (k & 63)) & 1 : 0; }
        } else if (op == OP_BITSLC) {
            const uint32_t d = code[p + 1], n = code[p + 2];
            p += 3;
            const fe v = eval_lc(code, &p, cm, W);
            for (uint32_t k = 0; k < n; k++) { f_zero(&W[d + k]); W[d + k].v[0] = k < 256 ? (v.v[k >> 6] >> (k & 63)) & 1 : 0; }
        } else if (op == OP_INV) {
            fe v = W[code[p + 2]], m;
            if (!fe_is_zero(&v)) { f_to_mont(&FR, &m, &v); f_inv(&FR, &m, &m); f_from_mont(&FR, &v, &m); }
            W[code[p + 1]] = v;
        } else { free(cm); return -1; }
    }
    free(cm);
    return failed ? -6 : 0;
}

/* ---- PLONK prover (oracle/plonk.py prove) ---- */
typedef struct { const uint8_t* p; uint64_t size; } section_t;
static int find_sections(const uint8_t* d, size_t len, const char* magic, section_t* secs, int max_id) {
    if (len < 12 || memcmp(d, magic, 4) != 0) return -1;
    uint32_t ns; memcpy(&ns, d + 8, 4);
    for (int i = 0; i <= max_id; i++) { secs[i].p = NULL; secs[i].size = 0; }
    size_t pos = 12;
    for (uint32_t i = 0; i < ns; i++) {
        uint32_t id; uint64_t sz; memcpy(&id, d + pos, 4); memcpy(&sz, d + pos + 4, 8); pos += 12;
        if (pos + sz > len) return -1;
        if ((int)id <= max_id && !secs[id].p) { secs[id].p = d + pos; secs[id].size = sz; }
        pos += sz;
    }
    return 0;
}
static void horner(fe* r, const fe* c, size_t n, const fe* x) {
    /* chunked so the threads share the work; exact field arithmetic => same value as the sequential loop */
    int nt = 1;
#ifdef _OPENMP
    nt = omp_get_max_threads();
#endif
    if (n < 4096) nt = 1;
    fe* part = (fe*)malloc(nt * sizeof(fe));
    size_t* lo = (size_t*)malloc((nt + 1) * sizeof(size_t));
    for (int t = 0; t <= nt; t++) lo[t] = n * (size_t)t / nt;
#pragma omp parallel for schedule(static, 1)
    for (int t = 0; t < nt; t++) {
        fe acc; f_zero(&acc);
        for (size_t i = lo[t + 1]; i > lo[t]; i--) { FRMUL(&acc, &acc, x); FRADD(&acc, &acc, &c[i - 1]); }
        part[t] = acc;
    }
    fe acc; f_zero(&acc);
    for (int t = nt - 1; t >= 0; t--) {
        /* acc = acc * x^(len of chunk t) + part[t] */
        uint64_t e[4] = {lo[t + 1] - lo[t], 0, 0, 0};
        fe xp; f_pow(&FR, &xp, x, e);
        FRMUL(&acc, &acc, &xp); FRADD(&acc, &acc, &part[t]);
    }
    *r = acc; free(part); free(lo);
}
static int div_pol1(fe* res, const fe* P, size_t n, const fe* d) {
    f_zero(&res[n - 1]);
    res[n - 2] = P[n - 1];
    for (size_t i = n - 2; i-- > 0;) { fe t; FRMUL(&t, d, &res[i + 1]); FRADD(&res[i], &P[i + 1], &t); }
    fe t, nd; f_neg(&FR, &nd, d); FRMUL(&t, &nd, &res[0]);
    return fe_eq(&P[0], &t) ? 0 : -5;
}
static void to4t(fe* pol, fe* ext, const fe* evals, size_t n, unsigned power, const fe* pz, int k) {
    memcpy(pol, evals, n * 32);
    ntt_core(pol, power, 1);
    memcpy(ext, pol, n * 32); memset(ext + n, 0, 3 * n * 32);
    ntt_core(ext, power + 2, 0);
    for (int i = 0; i < k; i++) { pol[n + i] = pz[i]; FRSUB(&pol[i], &pol[i], &pz[i]); }
}
typedef struct { fe r, rz; } pair_t;
static pair_t mul4(const fe* a, const fe* b, const fe* c, const fe* d, const fe* ap, const fe* bp, const fe* cp, const fe* dp,
                   const fe* z1, const fe* z2, const fe* z3) {
    fe a_b, a_bp, ap_b, ap_bp, c_d, c_dp, cp_d, cp_dp, a0, a1, a2, a3, t;
    FRMUL(&a_b, a, b); FRMUL(&a_bp, a, bp); FRMUL(&ap_b, ap, b); FRMUL(&ap_bp, ap, bp);
    FRMUL(&c_d, c, d); FRMUL(&c_dp, c, dp); FRMUL(&cp_d, cp, d); FRMUL(&cp_dp, cp, dp);
    pair_t o; FRMUL(&o.r, &a_b, &c_d);
    FRMUL(&a0, &ap_b, &c_d); FRMUL(&t, &a_bp, &c_d); FRADD(&a0, &a0, &t); FRMUL(&t, &a_b, &cp_d); FRADD(&a0, &a0, &t); FRMUL(&t, &a_b, &c_dp); FRADD(&a0, &a0, &t);
    FRMUL(&a1, &ap_bp, &c_d); FRMUL(&t, &ap_b, &cp_d); FRADD(&a1, &a1, &t); FRMUL(&t, &ap_b, &c_dp); FRADD(&a1, &a1, &t);
    FRMUL(&t, &a_bp, &cp_d); FRADD(&a1, &a1, &t); FRMUL(&t, &a_bp, &c_dp); FRADD(&a1, &a1, &t); FRMUL(&t, &a_b, &cp_dp); FRADD(&a1, &a1, &t);
    FRMUL(&a2, &a_bp, &cp_dp); FRMUL(&t, &ap_b, &cp_dp); FRADD(&a2, &a2, &t); FRMUL(&t, &ap_bp, &c_dp); FRADD(&a2, &a2, &t); FRMUL(&t, &ap_bp, &cp_d); FRADD(&a2, &a2, &t);
    FRMUL(&a3, &ap_bp, &cp_dp);
    o.rz = a0; FRMUL(&t, z1, &a1); FRADD(&o.rz, &o.rz, &t); FRMUL(&t, z2, &a2); FRADD(&o.rz, &o.rz, &t); FRMUL(&t, z3, &a3); FRADD(&o.rz, &o.rz, &t);
    return o;
}

/* proof_out: 800 bytes (nzcb_proof layout); public_out: nPublic x 32 LE.  w_le: n_w canonical values.
 * Returns 0, -3 (witness length), -4 (copy constraints), -5 (division). */
int oracle_prove_w(const uint8_t* zkey, size_t zlen, const uint8_t* w_le, uint32_t n_wit, const uint8_t* blinders_le,
                   uint8_t* proof_out, uint8_t* public_out) {
    section_t zs[15];
    if (find_sections(zkey, zlen, "zkey", zs, 14)) return -1;
    const uint8_t* h = zs[2].p;
    uint32_t f[5]; memcpy(f, h + 72, 20);
    const uint32_t n_vars = f[0], n_pub = f[1], n = f[2], n_add = f[3], n_cons = f[4];
    unsigned power = 0; while (((uint32_t)1 << power) < n) power++;
    if (n_wit != n_vars - n_add) return -3;
    fe k1, k2; memcpy(k1.v, h + 92, 32); memcpy(k2.v, h + 124, 32);
    const size_t N = n;
    fe bl[10];
    for (int i = 1; i <= 9; i++) { fe t; memcpy(t.v, blinders_le + (i - 1) * 32, 32); f_to_mont(&FR, &bl[i], &t); }
    /* witness -> Montgomery, w[0] := 0, additions */
    fe* W = (fe*)malloc((size_t)n_vars * sizeof(fe));
#pragma omp parallel for schedule(static)
    for (uint32_t i = 0; i < n_wit; i++) { fe t; memcpy(t.v, w_le + (size_t)i * 32, 32); f_to_mont(&FR, &W[i], &t); }
    f_zero(&W[0]);
    for (uint32_t i = 0; i < n_add; i++) {
        const uint8_t* a = zs[3].p + (size_t)i * 72;
        uint32_t ia, ib; memcpy(&ia, a, 4); memcpy(&ib, a + 4, 4);
        fe ac, bc, t1, t2, z; f_zero(&z); memcpy(ac.v, a + 8, 32); memcpy(bc.v, a + 40, 32);
        FRMUL(&t1, &ac, ia < n_vars ? &W[ia] : &z); FRMUL(&t2, &bc, ib < n_vars ? &W[ib] : &z);
        FRADD(&W[n_wit + i], &t1, &t2);
    }
    fe *A = (fe*)calloc(N, 32), *B = (fe*)calloc(N, 32), *C = (fe*)calloc(N, 32);
    for (uint32_t i = 0; i < n_cons; i++) {
        uint32_t s;
        memcpy(&s, zs[4].p + 4 * (size_t)i, 4); if (s < n_vars) A[i] = W[s];
        memcpy(&s, zs[5].p + 4 * (size_t)i, 4); if (s < n_vars) B[i] = W[s];
        memcpy(&s, zs[6].p + 4 * (size_t)i, 4); if (s < n_vars) C[i] = W[s];
    }
    free(W);
    const fe *QM = (const fe*)zs[7].p, *QL = (const fe*)zs[8].p, *QR = (const fe*)zs[9].p, *QO = (const fe*)zs[10].p, *QC = (const fe*)zs[11].p;
    const fe* SG = (const fe*)zs[12].p;
    const fe *S1 = SG, *S14 = SG + N, *S2 = SG + 5 * N, *S24 = SG + 6 * N, *S3 = SG + 10 * N, *S34 = SG + 11 * N;
    const fe* LG = (const fe*)zs[13].p;
    const g1a* PT = (const g1a*)zs[14].p;
    uint8_t* P = proof_out;
    g1a cm;
    /* round 1 */
    fe *pol_a = (fe*)malloc((N + 8) * 32), *pol_b = (fe*)malloc((N + 8) * 32), *pol_c = (fe*)malloc((N + 8) * 32), *pol_z = (fe*)malloc((N + 8) * 32);
    fe *A4 = (fe*)malloc(4 * N * 32), *B4 = (fe*)malloc(4 * N * 32), *C4 = (fe*)malloc(4 * N * 32), *Z4 = (fe*)malloc(4 * N * 32);
    { fe pz[2] = {bl[2], bl[1]}; to4t(pol_a, A4, A, N, power, pz, 2); }
    { fe pz[2] = {bl[4], bl[3]}; to4t(pol_b, B4, B, N, power, pz, 2); }
    { fe pz[2] = {bl[6], bl[5]}; to4t(pol_c, C4, C, N, power, pz, 2); }
    msm_core(&cm, PT, pol_a, N + 2, 1); g1_to_be(P + 0, &cm);
    msm_core(&cm, PT, pol_b, N + 2, 1); g1_to_be(P + 64, &cm);
    msm_core(&cm, PT, pol_c, N + 2, 1); g1_to_be(P + 128, &cm);
    /* round 2 */
    uint8_t tr[64 * 9 + 32 * 64]; size_t tl = 0;
    for (uint32_t i = 0; i < n_pub; i++) { fr_to_be(tr + tl, &A[i]); tl += 32; if (public_out) memcpy(public_out + 32 * (size_t)i, w_le + 32 * (size_t)(i + 1), 32); }
    memcpy(tr + tl, P, 192); tl += 192;
    fe beta, gamma; hash_to_fr(&beta, tr, tl);
    { uint8_t b32[32]; fr_to_be(b32, &beta); hash_to_fr(&gamma, b32, 32); }
    fe wn; fr_root(&wn, power);
    fe* Wn = (fe*)malloc(N * 32);
    f_one(&FR, &Wn[0]); for (size_t i = 1; i < N; i++) FRMUL(&Wn[i], &Wn[i - 1], &wn);
    fe *ratio = (fe*)malloc(N * 32), *Z = (fe*)malloc(N * 32);
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < N; i++) {
        fe bw, n1, n2, n3, d1, d2, d3, t, num, den;
        FRMUL(&bw, &beta, &Wn[i]);
        FRADD(&n1, &A[i], &bw); FRADD(&n1, &n1, &gamma);
        FRMUL(&t, &k1, &bw); FRADD(&n2, &B[i], &t); FRADD(&n2, &n2, &gamma);
        FRMUL(&t, &k2, &bw); FRADD(&n3, &C[i], &t); FRADD(&n3, &n3, &gamma);
        FRMUL(&num, &n2, &n3); FRMUL(&num, &n1, &num);
        FRMUL(&t, &S14[4 * i], &beta); FRADD(&d1, &A[i], &t); FRADD(&d1, &d1, &gamma);
        FRMUL(&t, &S24[4 * i], &beta); FRADD(&d2, &B[i], &t); FRADD(&d2, &d2, &gamma);
        FRMUL(&t, &S34[4 * i], &beta); FRADD(&d3, &C[i], &t); FRADD(&d3, &d3, &gamma);
        FRMUL(&den, &d2, &d3); FRMUL(&den, &d1, &den);
        ratio[i] = num; Z[i] = den;
    }
    /* batchInverse of the denominators (Montgomery's trick per chunk), then ratio = num / den */
#pragma omp parallel for schedule(static)
    for (size_t c0 = 0; c0 < N; c0 += 1024) {
        const size_t c1 = c0 + 1024 < N ? c0 + 1024 : N;
        fe pre[1024], acc, inv;
        acc = Z[c0]; pre[0] = acc;
        for (size_t i = c0 + 1; i < c1; i++) { FRMUL(&acc, &acc, &Z[i]); pre[i - c0] = acc; }
        f_inv(&FR, &inv, &acc);
        for (size_t i = c1 - 1; i > c0; i--) { fe d = Z[i], t; FRMUL(&t, &inv, &pre[i - 1 - c0]); FRMUL(&inv, &inv, &d); FRMUL(&ratio[i], &ratio[i], &t); }
        FRMUL(&ratio[c0], &ratio[c0], &inv);
    }
    f_one(&FR, &Z[0]);
    for (size_t i = 1; i < N; i++) FRMUL(&Z[i], &Z[i - 1], &ratio[i - 1]);
    { fe last, one; FRMUL(&last, &Z[N - 1], &ratio[N - 1]); f_one(&FR, &one); if (!fe_eq(&last, &one)) return -4; }
    free(ratio);
    { fe pz[3] = {bl[9], bl[8], bl[7]}; to4t(pol_z, Z4, Z, N, power, pz, 3); }
    msm_core(&cm, PT, pol_z, N + 3, 1); g1_to_be(P + 192, &cm);
    /* round 3 */
    fe alpha, alpha2; hash_to_fr(&alpha, P + 192, 64); FRMUL(&alpha2, &alpha, &alpha);
    fe w4, one, two, four, eight, zero; fr_root(&w4, 2); f_one(&FR, &one); f_zero(&zero);
    FRADD(&two, &one, &one); FRADD(&four, &two, &two); FRADD(&eight, &four, &four);
    fe Z1[4], Z2[4], Z3[4], t0;
    Z1[0] = zero; FRSUB(&Z1[1], &w4, &one); FRSUB(&Z1[2], &zero, &two); FRSUB(&Z1[3], &zero, &one); FRSUB(&Z1[3], &Z1[3], &w4);
    FRMUL(&t0, &two, &w4); Z2[0] = zero; FRSUB(&Z2[1], &zero, &t0); Z2[2] = four; Z2[3] = t0;
    Z3[0] = zero; FRADD(&Z3[1], &two, &t0); FRSUB(&Z3[2], &zero, &eight); FRSUB(&Z3[3], &two, &t0);
    fe w4n; fr_root(&w4n, power + 2);
    fe* X4 = (fe*)malloc(4 * N * 32);
    f_one(&FR, &X4[0]); for (size_t i = 1; i < 4 * N; i++) FRMUL(&X4[i], &X4[i - 1], &w4n);
    fe *T = (fe*)malloc(4 * N * 32), *Tz = (fe*)malloc(4 * N * 32);
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < 4 * N; i++) {
        const int p = (int)(i & 3);
        const fe x = X4[i];
        const fe *a = &A4[i], *b = &B4[i], *c = &C4[i], *z = &Z4[i], *zw = &Z4[(i + 4) & (4 * N - 1)];
        fe ap, bp, cp, zp, zwp, xw, t, pl, e1, e1z, bx, ea, eb, ec, e4, e4z, l1 = LG[N + i];
        FRMUL(&t, &bl[1], &x); FRADD(&ap, &bl[2], &t);
        FRMUL(&t, &bl[3], &x); FRADD(&bp, &bl[4], &t);
        FRMUL(&t, &bl[5], &x); FRADD(&cp, &bl[6], &t);
        FRMUL(&t, &bl[7], &x); FRADD(&t, &t, &bl[8]); FRMUL(&t, &t, &x); FRADD(&zp, &t, &bl[9]);
        FRMUL(&xw, &x, &wn);
        FRMUL(&t, &bl[7], &xw); FRADD(&t, &t, &bl[8]); FRMUL(&t, &t, &xw); FRADD(&zwp, &t, &bl[9]);
        f_zero(&pl);
        for (uint32_t j = 0; j < n_pub; j++) { FRMUL(&t, &LG[(size_t)j * 5 * N + N + i], &A[j]); FRSUB(&pl, &pl, &t); }
        /* e1 */
        fe ab, abp, apb, apbp, rz;
        FRMUL(&ab, a, b); FRMUL(&abp, a, &bp); FRMUL(&apb, &ap, b); FRMUL(&apbp, &ap, &bp);
        FRADD(&rz, &abp, &apb); FRMUL(&t, &Z1[p], &apbp); FRADD(&rz, &rz, &t);
        FRMUL(&e1, &ab, &QM[N + i]); FRMUL(&t, a, &QL[N + i]); FRADD(&e1, &e1, &t); FRMUL(&t, b, &QR[N + i]); FRADD(&e1, &e1, &t);
        FRMUL(&t, c, &QO[N + i]); FRADD(&e1, &e1, &t); FRADD(&e1, &e1, &pl); FRADD(&e1, &e1, &QC[N + i]);
        FRMUL(&e1z, &rz, &QM[N + i]); FRMUL(&t, &ap, &QL[N + i]); FRADD(&e1z, &e1z, &t); FRMUL(&t, &bp, &QR[N + i]); FRADD(&e1z, &e1z, &t);
        FRMUL(&t, &cp, &QO[N + i]); FRADD(&e1z, &e1z, &t);
        /* e2, e3 */
        FRMUL(&bx, &beta, &x);
        FRADD(&ea, a, &bx); FRADD(&ea, &ea, &gamma);
        FRMUL(&t, &bx, &k1); FRADD(&eb, b, &t); FRADD(&eb, &eb, &gamma);
        FRMUL(&t, &bx, &k2); FRADD(&ec, c, &t); FRADD(&ec, &ec, &gamma);
        pair_t e2 = mul4(&ea, &eb, &ec, z, &ap, &bp, &cp, &zp, &Z1[p], &Z2[p], &Z3[p]);
        FRMUL(&t, &beta, &S14[i]); FRADD(&ea, a, &t); FRADD(&ea, &ea, &gamma);
        FRMUL(&t, &beta, &S24[i]); FRADD(&eb, b, &t); FRADD(&eb, &eb, &gamma);
        FRMUL(&t, &beta, &S34[i]); FRADD(&ec, c, &t); FRADD(&ec, &ec, &gamma);
        pair_t e3 = mul4(&ea, &eb, &ec, zw, &ap, &bp, &cp, &zwp, &Z1[p], &Z2[p], &Z3[p]);
        FRSUB(&t, z, &one); FRMUL(&e4, &t, &l1); FRMUL(&e4, &e4, &alpha2);
        FRMUL(&e4z, &zp, &l1); FRMUL(&e4z, &e4z, &alpha2);
        fe d; FRSUB(&d, &e2.r, &e3.r); FRMUL(&d, &d, &alpha); FRADD(&T[i], &e1, &d); FRADD(&T[i], &T[i], &e4);
        FRSUB(&d, &e2.rz, &e3.rz); FRMUL(&d, &d, &alpha); FRADD(&Tz[i], &e1z, &d); FRADD(&Tz[i], &Tz[i], &e4z);
    }
    free(X4); free(A4); free(B4); free(C4); free(Z4);
    ntt_core(T, power + 2, 1);
    int bad = 0;
#pragma omp parallel for schedule(static) reduction(|:bad)
    for (size_t i = 0; i < N; i++) {
        fe prev; f_neg(&FR, &prev, &T[i]); T[i] = prev;
        for (int k = 1; k < 4; k++) { size_t idx = k * N + i; FRSUB(&prev, &prev, &T[idx]); T[idx] = prev; if (idx > 3 * N - 4 && !fe_is_zero(&prev)) bad |= 1; }
    }
    if (bad) return -5;
    ntt_core(Tz, power + 2, 1);
    for (size_t i = 0; i < 4 * N; i++) { if (i > 3 * N + 5) { if (!fe_is_zero(&Tz[i])) return -5; } else FRADD(&T[i], &T[i], &Tz[i]); }
    free(Tz);
    fe* pol_t = T;
    msm_core(&cm, PT, pol_t, N, 1); g1_to_be(P + 256, &cm);
    msm_core(&cm, PT, pol_t + N, N, 1); g1_to_be(P + 320, &cm);
    msm_core(&cm, PT, pol_t + 2 * N, N + 6, 1); g1_to_be(P + 384, &cm);
    /* round 4 */
    fe xi, xiw; hash_to_fr(&xi, P + 256, 192); FRMUL(&xiw, &xi, &wn);
    fe ev_a, ev_b, ev_c, ev_s1, ev_s2, ev_t, ev_zw, ev_r;
    horner(&ev_a, pol_a, N + 2, &xi); horner(&ev_b, pol_b, N + 2, &xi); horner(&ev_c, pol_c, N + 2, &xi);
    horner(&ev_s1, S1, N, &xi); horner(&ev_s2, S2, N, &xi); horner(&ev_t, pol_t, 3 * N + 6, &xi); horner(&ev_zw, pol_z, N + 3, &xiw);
    fe coef_ab, bxi, e2, e3, t, u, xim, l1, e4, coefz;
    FRMUL(&coef_ab, &ev_a, &ev_b); FRMUL(&bxi, &beta, &xi);
    FRADD(&e2, &ev_a, &bxi); FRADD(&e2, &e2, &gamma);
    FRMUL(&t, &bxi, &k1); FRADD(&u, &ev_b, &t); FRADD(&u, &u, &gamma); FRMUL(&e2, &e2, &u);
    FRMUL(&t, &bxi, &k2); FRADD(&u, &ev_c, &t); FRADD(&u, &u, &gamma); FRMUL(&e2, &e2, &u); FRMUL(&e2, &e2, &alpha);
    FRMUL(&t, &beta, &ev_s1); FRADD(&e3, &ev_a, &t); FRADD(&e3, &e3, &gamma);
    FRMUL(&t, &beta, &ev_s2); FRADD(&u, &ev_b, &t); FRADD(&u, &u, &gamma); FRMUL(&e3, &e3, &u);
    FRMUL(&e3, &e3, &beta); FRMUL(&e3, &e3, &ev_zw); FRMUL(&e3, &e3, &alpha);
    xim = xi; for (unsigned i = 0; i < power; i++) f_sqr(&FR, &xim, &xim);
    { fe nn, d; f_from_u64(&FR, &nn, n); FRSUB(&d, &xi, &one); FRMUL(&d, &d, &nn); f_inv(&FR, &d, &d); FRSUB(&l1, &xim, &one); FRMUL(&l1, &l1, &d); }
    FRMUL(&e4, &l1, &alpha2); FRADD(&coefz, &e2, &e4);
    fe* pol_r = (fe*)malloc((N + 8) * 32);
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < N + 3; i++) {
        fe v, q; FRMUL(&v, &coefz, &pol_z[i]);
        if (i < N) {
            FRMUL(&q, &coef_ab, &QM[i]); FRADD(&v, &v, &q); FRMUL(&q, &ev_a, &QL[i]); FRADD(&v, &v, &q);
            FRMUL(&q, &ev_b, &QR[i]); FRADD(&v, &v, &q); FRMUL(&q, &ev_c, &QO[i]); FRADD(&v, &v, &q);
            FRADD(&v, &v, &QC[i]); FRMUL(&q, &e3, &S3[i]); FRSUB(&v, &v, &q);
        }
        pol_r[i] = v;
    }
    horner(&ev_r, pol_r, N + 3, &xi);
    fr_to_be(P + 576, &ev_a); fr_to_be(P + 608, &ev_b); fr_to_be(P + 640, &ev_c); fr_to_be(P + 672, &ev_s1);
    fr_to_be(P + 704, &ev_s2); fr_to_be(P + 736, &ev_zw); fr_to_be(P + 768, &ev_r);
    /* round 5 */
    fe v[7]; v[0] = one; hash_to_fr(&v[1], P + 576, 224);
    for (int i = 2; i <= 6; i++) FRMUL(&v[i], &v[i - 1], &v[1]);
    fe xi2m; FRMUL(&xi2m, &xim, &xim);
    fe *wxi = (fe*)malloc((N + 8) * 32), *quot = (fe*)malloc((N + 8) * 32);
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < N + 6; i++) {
        fe w, q; FRMUL(&w, &xi2m, &pol_t[2 * N + i]);
        if (i < N + 3) { FRMUL(&q, &v[1], &pol_r[i]); FRADD(&w, &w, &q); }
        if (i < N + 2) { FRMUL(&q, &v[2], &pol_a[i]); FRADD(&w, &w, &q); FRMUL(&q, &v[3], &pol_b[i]); FRADD(&w, &w, &q); FRMUL(&q, &v[4], &pol_c[i]); FRADD(&w, &w, &q); }
        if (i < N) { FRADD(&w, &w, &pol_t[i]); FRMUL(&q, &xim, &pol_t[N + i]); FRADD(&w, &w, &q); FRMUL(&q, &v[5], &S1[i]); FRADD(&w, &w, &q); FRMUL(&q, &v[6], &S2[i]); FRADD(&w, &w, &q); }
        wxi[i] = w;
    }
    { fe s = ev_t, q; FRMUL(&q, &v[1], &ev_r); FRADD(&s, &s, &q); FRMUL(&q, &v[2], &ev_a); FRADD(&s, &s, &q); FRMUL(&q, &v[3], &ev_b); FRADD(&s, &s, &q);
      FRMUL(&q, &v[4], &ev_c); FRADD(&s, &s, &q); FRMUL(&q, &v[5], &ev_s1); FRADD(&s, &s, &q); FRMUL(&q, &v[6], &ev_s2); FRADD(&s, &s, &q); FRSUB(&wxi[0], &wxi[0], &s); }
    if (div_pol1(quot, wxi, N + 6, &xi)) return -5;
    msm_core(&cm, PT, quot, N + 6, 1); g1_to_be(P + 448, &cm);
    memcpy(wxi, pol_z, (N + 3) * 32); FRSUB(&wxi[0], &wxi[0], &ev_zw);
    if (div_pol1(quot, wxi, N + 3, &xiw)) return -5;
    msm_core(&cm, PT, quot, N + 3, 1); g1_to_be(P + 512, &cm);
    free(A); free(B); free(C); free(Wn); free(Z); free(pol_a); free(pol_b); free(pol_c); free(pol_z); free(T); free(pol_r); free(wxi); free(quot);
    return 0;
}

/* plonk.prove(zkey, wtns) */
int oracle_prove(const uint8_t* zkey, size_t zlen, const uint8_t* wtns, size_t wlen, const uint8_t* blinders_le,
                 uint8_t* proof_out, uint8_t* public_out) {
    section_t ws[3];
    if (find_sections(wtns, wlen, "wtns", ws, 2) || !ws[1].p || !ws[2].p) return -1;
    uint32_t nw; memcpy(&nw, ws[1].p + 36, 4);
    if (ws[2].size != (uint64_t)nw * 32) return -3;
    return oracle_prove_w(zkey, zlen, ws[2].p, nw, blinders_le, proof_out, public_out);
}

/* plonk.fullProve: witness program then prover, all on the CPU */
int oracle_fullprove(const uint8_t* wprog, size_t plen, const uint8_t* inputs_le, const uint8_t* zkey, size_t zlen,
                     const uint8_t* blinders_le, uint8_t* proof_out, uint8_t* public_out) {
    if (plen < 40) return -1;
    uint32_t h[9]; memcpy(h, wprog + 4, 36);
    uint8_t* wires = (uint8_t*)malloc((size_t)h[1] * 32);
    int rc = oracle_witness(wprog, plen, inputs_le, wires);
    if (rc == 0) rc = oracle_prove_w(zkey, zlen, wires, h[2], blinders_le, proof_out, public_out);
    free(wires);
    return rc;
}

int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* torchrun exports OMP_NUM_THREADS=1 to its ranks: the baseline legs ask for the host's cores explicitly */
void oracle_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}
