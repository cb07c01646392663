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
