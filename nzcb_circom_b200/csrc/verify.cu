// plonk.verify on the device: snarkjs 0.4.12 plonk.verify(vk, publicSignals, proof) with the BN254
// pairing check curve.pairingEq(-A1, X_2, B1, G2.one), the verification key from the zkey header or
// from verification_key.json (`snarkjs zkey export verificationkey`, /root/reference/Makefile:56,61), [tau]_2 for the
// synthetic SRS, and `snarkjs zkey export soliditycalldata`'s text for a proof.
//
// One warp per proof in the latency form: lane 0 checks the proof's form, replays the Fiat-Shamir transcript
// (Keccak-256) and derives the twenty scalars; lanes 0..19 each do one G1 scalar multiplication; lanes 0 and 1 add up
// B1 and A1 and run one Miller loop each; lane 0 multiplies the two and does the final exponentiation.  A
// verification is latency-bound (a chain of ~25 k dependent Fq multiplications), so throughput comes from the batch:
// B proofs = B warps, and once the GPU's warp slots are full, up to 16 proofs per warp (see k_plonk_verify).
#include <stdlib.h>
#include "common.cuh"
#include "verify.cuh"

using namespace nzcb;

struct nzcb_vkey {
    nzcb_ctx* ctx = nullptr;
    VkDev h;              // host copy
    VkDev* d = nullptr;   // device copy
};

namespace {

// One warp verifies G proofs at a time (G = 1, 2, 4, 8 or 16, chosen by the host from the batch size).  Every phase
// has ONE call site that all participating lanes reach together, so the lanes run side by side:
//   A  lanes < G        : form checks, transcript, scalars of "their" proof            (verify_prepare)
//   B  20 G work items  : one G1 scalar multiplication each, 32 at a time              (g1_mul_limbs)
//   C  lanes < 2 G      : B1 and -A1 of proof lane/2                                   (g1_sum_affine)
//   D  lanes < 2 G      : one Miller loop each                                         (miller_loop)
//   E  lanes < G        : product of the two loops, final exponentiation, verdict     (final_exp)
// G = 1 is the latency form (one proof spread over a warp: 20 lanes in B, 2 in D); G = 16 fills the lanes in every
// phase but A and E (half) -- the throughput form for large batches.  serial = 1 (NZCB_VERIFY_SERIAL=1, blockDim 1)
// walks the lane index on one thread: a cross-check of the lane choreography.
constexpr uint32_t VERIFY_GMAX = 16;
// Shared memory per proof of the group, 2.6 KB: a work item's point and scalar (96 B, phase A -> B) are overwritten by
// its product (128 B, phase B -> C); the two summed points and then the two Miller values (phase C -> D -> E) reuse
// the items' space once the products are dead.  16 proofs = 41.5 KB, so five groups fit an SM.
struct VerifyItem {
    union {
        struct {
            G1Affine pt;
            Fr sc;
        } in;
        G1XYZZ acc;
    };
};
struct VerifySmem {
    VerifyItem item[VERIFY_TERMS];
    int ok, pad[7];
};
static_assert(sizeof(VerifyItem) == 128 && 2 * sizeof(Fq12) + 2 * sizeof(G1Affine) <= 8 * sizeof(VerifyItem), "layout");

__global__ void __launch_bounds__(32) k_plonk_verify(const VkDev* __restrict__ vkp, const uint8_t* __restrict__ proofs,
                                                     const uint8_t* __restrict__ pubs, uint32_t n_pub, uint32_t B,
                                                     int32_t* __restrict__ valid, int serial, uint32_t G) {
    extern __shared__ uint4 verify_smem_raw[];
    VerifySmem* sm = reinterpret_cast<VerifySmem*>(verify_smem_raw);
    const uint32_t n_groups = (B + G - 1) / G;
    const uint32_t width = serial ? 1u : 32u;  // lanes that walk the work items of a phase
    G1Affine serial_first = G1Affine::inf();
    for (uint32_t grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
        const uint32_t b0 = grp * G, g_here = min(G, B - b0);
        // A  (verify_prepare fills plain arrays: staged in this lane's stack, then scattered into the items)
#pragma unroll 1
        for (uint32_t w = threadIdx.x; w < g_here; w += width) {
            const uint32_t b = b0 + w;
            G1Affine pts[VERIFY_TERMS];
            Fr sc[VERIFY_TERMS];
            const bool ok = verify_prepare(*vkp, proofs + (size_t)b * sizeof(nzcb_proof), pubs + (size_t)b * n_pub * 32, n_pub,
                                           pts, sc) && !vkp->X2.is_inf() && g2_on_curve(vkp->X2);
            sm[w].ok = ok;
            if (ok)
                for (int t = 0; t < VERIFY_TERMS; t++) {
                    sm[w].item[t].in.pt = pts[t];
                    sm[w].item[t].in.sc = sc[t];
                }
        }
        __syncwarp();
        // B
#pragma unroll 1
        for (uint32_t w = threadIdx.x; w < g_here * VERIFY_TERMS; w += width) {
            const uint32_t p = w / VERIFY_TERMS, t = w % VERIFY_TERMS;
            if (sm[p].ok) {
                const G1Affine P = sm[p].item[t].in.pt;
                const Fr k = sm[p].item[t].in.sc;
                sm[p].item[t].acc = g1_mul_limbs(P, k);
            }
        }
        __syncwarp();
        // C: both sums of a proof are taken before either is stored (they overwrite products 0 and 1)
        G1Affine sum = G1Affine::inf();
#pragma unroll 1
        for (uint32_t w = threadIdx.x; w < g_here * 2; w += width) {
            const uint32_t p = w >> 1, l = w & 1;
            if (sm[p].ok) sum = g1_sum_affine(l == 0 ? &sm[p].item[0].acc : &sm[p].item[18].acc, l == 0 ? 18 : 2, l == 1);
            if (serial) {  // one thread: keep the first sum in a register-resident copy until the second is done
                if (l == 0) serial_first = sum;
                else if (sm[p].ok) {
                    G1Affine* ab = reinterpret_cast<G1Affine*>(&sm[p].item[0]);
                    ab[1] = serial_first;
                    ab[0] = sum;
                }
            }
        }
        __syncwarp();
        if (!serial) {
            const uint32_t w = threadIdx.x;
            if (w < g_here * 2 && sm[w >> 1].ok) reinterpret_cast<G1Affine*>(&sm[w >> 1].item[0])[1 - (w & 1)] = sum;
        }
        __syncwarp();
        // D: the Miller values go to items 2..7 (ab lives in item 0)
#pragma unroll 1
        for (uint32_t w = threadIdx.x; w < g_here * 2; w += width) {
            const uint32_t p = w >> 1, l = w & 1;
            if (sm[p].ok) {
                const G2Affine q = l == 0 ? vkp->X2 : g2_generator();
                const G1Affine a = reinterpret_cast<const G1Affine*>(&sm[p].item[0])[l];
                reinterpret_cast<Fq12*>(&sm[p].item[2])[l] = miller_loop(a, q);
            }
        }
        __syncwarp();
        // E
#pragma unroll 1
        for (uint32_t w = threadIdx.x; w < g_here; w += width) {
            const Fq12* fm = reinterpret_cast<const Fq12*>(&sm[w].item[2]);
            valid[b0 + w] = sm[w].ok && final_exp(f12_mul(fm[0], fm[1])).is_one() ? 1 : 0;
        }
        __syncwarp();
    }
}

// curve.pairingEq: prod e(P_i, Q_i) == 1; one thread per pair for the Miller loops, thread 0 finishes
__global__ void __launch_bounds__(32) k_pairing_eq(const G1Affine* __restrict__ P, const G2Affine* __restrict__ Q, uint32_t n,
                                                   int32_t* __restrict__ result, Fq12* __restrict__ gt_out) {
    __shared__ Fq12 f[32];
    __shared__ int bad;
    if (threadIdx.x == 0) bad = 0;
    __syncwarp();
    if (threadIdx.x < n) {
        const G1Affine p = P[threadIdx.x];
        const G2Affine q = Q[threadIdx.x];
        const bool on = (p.is_inf() || p.y.sqr() == p.x.sqr() * p.x + Fq::from_u64(3)) && g2_on_curve(q);
        if (!on) atomicExch(&bad, 1);
        f[threadIdx.x] = on ? miller_loop(p, q) : Fq12::one();
    }
    __syncwarp();
    if (threadIdx.x == 0) {
        Fq12 t = Fq12::one();
        for (uint32_t i = 0; i < n; i++) t = f12_mul(t, f[i]);
        t = final_exp(t);
        if (gt_out) *gt_out = t;
        *result = bad ? -1 : (t.is_one() ? 1 : 0);
    }
}

__global__ void k_g2_mul(G2Affine q, Fr k_canonical, G2Affine* out) { *out = g2_mul_limbs(q, k_canonical.v); }

// ---- verification_key.json: decimal strings -> limbs -------------------------------------------------------
bool dec_to_limbs(const char* s, size_t n, uint32_t out[8]) {
    for (int i = 0; i < 8; i++) out[i] = 0;
    if (n == 0) return false;
    for (size_t k = 0; k < n; k++) {
        if (s[k] < '0' || s[k] > '9') return false;
        uint64_t carry = (uint64_t)(s[k] - '0');
        for (int i = 0; i < 8; i++) {
            const uint64_t cur = (uint64_t)out[i] * 10 + carry;
            out[i] = (uint32_t)cur;
            carry = cur >> 32;
        }
        if (carry) return false;
    }
    return true;
}
// the decimal numbers (quoted or bare) that follow "key": up to the matching close of its value, at most `max`
int json_numbers(const std::string& js, const char* key, uint32_t (*out)[8], int max) {
    const std::string pat = std::string("\"") + key + "\"";
    size_t p = js.find(pat);
    if (p == std::string::npos) return -1;
    p = js.find(':', p + pat.size());
    if (p == std::string::npos) return -1;
    int depth = 0, got = 0;
    for (size_t i = p + 1; i < js.size(); i++) {
        const char c = js[i];
        if (c == '[') depth++;
        else if (c == ']') {
            if (--depth <= 0) break;
        } else if (c >= '0' && c <= '9') {
            size_t e = i;
            while (e < js.size() && js[e] >= '0' && js[e] <= '9') e++;
            if (got >= max || !dec_to_limbs(js.data() + i, e - i, out[got])) return -1;
            got++;
            i = e - 1;
            if (depth == 0) {  // a scalar value: done after the first number
                break;
            }
        } else if ((c == ',' || c == '}') && depth == 0 && got) break;
    }
    return got;
}
template <class F>
bool limbs_to_mont(const uint32_t w[8], F& out) {
    F x;
    for (int i = 0; i < 8; i++) x.v[i] = w[i];
    const F m = F::modulus();
    bool lt = false;
    for (int i = 7; i >= 0; i--) {
        if (x.v[i] != m.v[i]) {
            lt = x.v[i] < m.v[i];
            break;
        }
    }
    if (!lt) return false;
    out = x.to_mont();
    return true;
}

int vkey_finish(nzcb_ctx* ctx, nzcb_vkey* vk, nzcb_vkey** out) {
    vk->ctx = ctx;
    if (vk->h.power > 28) {
        delete vk;
        return ctx->fail(NZCB_E_INVALID, "verification key: power %u out of range", vk->h.power);
    }
    // X_2 = [tau]_2 must be a point of order r.  With X_2 at infinity e(., X_2) = 1 and the pairing check no longer
    // binds the proof to the key (a forged proof with Z = Wxiw = infinity passes); a point of the twist outside the
    // r-torsion would leave the pairing undefined.  Keys written without X_2 (all zero) can prove, not verify.
    {
        uint32_t r1[8];  // r - 1 (r is odd): [r - 1] X_2 == -X_2  <=>  [r] X_2 == infinity
        for (int i = 0; i < 8; i++) r1[i] = Fr::modulus().v[i];
        r1[0] -= 1;
        bool ok = !vk->h.X2.is_inf() && g2_on_curve(vk->h.X2);
        if (ok) {
            const G2Affine m = g2_mul_limbs(vk->h.X2, r1);
            ok = !m.is_inf() && m.x == vk->h.X2.x && m.y == vk->h.X2.y.neg();
        }
        if (!ok) {
            delete vk;
            return ctx->fail(NZCB_E_INVALID, "verification key: X_2 is not a point of order r (missing or invalid [tau]_2)");
        }
    }
    // w = 5^((r-1)/2^28) squared down to the 2^power-th root
    Fr w = Fr::from_u64(5);
    {
        uint32_t e[8];
        const Fr m = Fr::modulus();
        for (int i = 0; i < 8; i++) e[i] = m.v[i];
        e[0] -= 1;
        // (r - 1) >> 28
        uint32_t sh[8];
        for (int i = 0; i < 8; i++) sh[i] = (e[i] >> 28) | (i + 1 < 8 ? e[i + 1] << 4 : 0);
        w = w.pow_limbs(sh);
    }
    for (uint32_t i = vk->h.power; i < 28; i++) w = w.sqr();
    vk->h.w = w;
    cudaSetDevice(ctx->device);
    if (cudaMalloc(&vk->d, sizeof(VkDev)) != cudaSuccess ||
        cudaMemcpy(vk->d, &vk->h, sizeof(VkDev), cudaMemcpyHostToDevice) != cudaSuccess) {
        cudaGetLastError();
        if (vk->d) cudaFree(vk->d);
        delete vk;
        return ctx->fail(NZCB_E_CUDA, "verification key: cannot upload");
    }
    *out = vk;
    return 0;
}
}  // namespace

extern "C" int32_t nzcb_vkey_from_zkey(nzcb_ctx* ctx, const uint8_t* zkey, size_t len, nzcb_vkey** out) {
    if (!ctx || !zkey || !out) return NZCB_E_INVALID;
    *out = nullptr;
    if (len < 12 || memcmp(zkey, "zkey", 4) != 0) return ctx->fail(NZCB_E_INVALID, "zkey file: bad magic");
    uint32_t nsec;
    memcpy(&nsec, zkey + 8, 4);
    size_t pos = 12;
    const uint8_t* hdr = nullptr;
    uint64_t hsize = 0;
    for (uint32_t s = 0; s < nsec && pos + 12 <= len; s++) {
        uint32_t id;
        uint64_t size;
        memcpy(&id, zkey + pos, 4);
        memcpy(&size, zkey + pos + 4, 8);
        pos += 12;
        if (size > len - pos) return ctx->fail(NZCB_E_INVALID, "zkey file: section %u overruns the file", id);
        if (id == 2) {
            hdr = zkey + pos;
            hsize = size;
            break;
        }
        pos += size;
    }
    // header: n8q, q, n8r, r, nVars, nPublic, domainSize, nAdditions, nConstraints, k1, k2, 8 G1, X_2
    if (!hdr || hsize < 156 + 8 * 64 + 128) return ctx->fail(NZCB_E_INVALID, "zkey file: no PLONK header section");
    nzcb_vkey* vk = new nzcb_vkey();
    uint32_t domain;
    memcpy(&vk->h.n_public, hdr + 76, 4);
    memcpy(&domain, hdr + 80, 4);
    if (domain == 0 || (domain & (domain - 1))) {
        delete vk;
        return ctx->fail(NZCB_E_INVALID, "zkey file: domain size %u is not a power of two", domain);
    }
    vk->h.power = 0;
    while ((1u << vk->h.power) < domain) vk->h.power++;
    memcpy(&vk->h.k1, hdr + 92, 32);
    memcpy(&vk->h.k2, hdr + 124, 32);
    memcpy(vk->h.Q, hdr + 156, 512);
    memcpy(&vk->h.X2, hdr + 156 + 512, 128);
    return vkey_finish(ctx, vk, out);
}

extern "C" int32_t nzcb_vkey_from_json(nzcb_ctx* ctx, const char* json, size_t len, nzcb_vkey** out) {
    if (!ctx || !json || !out) return NZCB_E_INVALID;
    *out = nullptr;
    const std::string js(json, len);
    if (js.find("\"plonk\"") == std::string::npos) return ctx->fail(NZCB_E_INVALID, "verification key: protocol is not plonk");
    nzcb_vkey* vk = new nzcb_vkey();
    uint32_t num[6][8];
    auto bad = [&](const char* what) {
        delete vk;
        return ctx->fail(NZCB_E_INVALID, "verification key: bad or missing \"%s\"", what);
    };
    if (json_numbers(js, "nPublic", num, 1) != 1) return bad("nPublic");
    vk->h.n_public = num[0][0];
    if (json_numbers(js, "power", num, 1) != 1) return bad("power");
    vk->h.power = num[0][0];
    if (json_numbers(js, "k1", num, 1) != 1 || !limbs_to_mont(num[0], vk->h.k1)) return bad("k1");
    if (json_numbers(js, "k2", num, 1) != 1 || !limbs_to_mont(num[0], vk->h.k2)) return bad("k2");
    static const char* names[8] = {"Qm", "Ql", "Qr", "Qo", "Qc", "S1", "S2", "S3"};
    for (int i = 0; i < 8; i++) {
        if (json_numbers(js, names[i], num, 3) != 3) return bad(names[i]);
        bool z_zero = true;
        for (int k = 0; k < 8; k++) z_zero = z_zero && num[2][k] == 0;
        if (z_zero) vk->h.Q[i] = G1Affine::inf();  // [x, y, "0"]: infinity
        else if (!limbs_to_mont(num[0], vk->h.Q[i].x) || !limbs_to_mont(num[1], vk->h.Q[i].y)) return bad(names[i]);
    }
    if (json_numbers(js, "X_2", num, 6) != 6) return bad("X_2");
    bool z_zero = true;
    for (int k = 0; k < 8; k++) z_zero = z_zero && num[4][k] == 0 && num[5][k] == 0;
    if (z_zero) vk->h.X2 = G2Affine{Fq2::zero(), Fq2::zero()};
    else if (!limbs_to_mont(num[0], vk->h.X2.x.c0) || !limbs_to_mont(num[1], vk->h.X2.x.c1) ||
             !limbs_to_mont(num[2], vk->h.X2.y.c0) || !limbs_to_mont(num[3], vk->h.X2.y.c1))
        return bad("X_2");
    return vkey_finish(ctx, vk, out);
}

extern "C" void nzcb_vkey_free(nzcb_vkey* vk) {
    if (!vk) return;
    if (vk->d) {
        cudaSetDevice(vk->ctx->device);
        cudaFree(vk->d);
    }
    delete vk;
}

extern "C" int32_t nzcb_plonk_verify_batch(nzcb_ctx* ctx, const nzcb_vkey* vk, const nzcb_proof* proofs,
                                           const uint8_t* public_le, uint32_t n_public, size_t B, int32_t* valid) {
    if (!ctx || !vk || !proofs || !valid || (n_public && !public_le)) return NZCB_E_INVALID;
    if (vk->ctx != ctx) return ctx->fail(NZCB_E_INVALID, "verification key was loaded on a different context");
    if (B == 0) return 0;
    NZ_CUDA(ctx, cudaSetDevice(ctx->device));
    uint8_t* d_proofs = (uint8_t*)ctx->scratch_get("vf_proofs", B * sizeof(nzcb_proof));
    uint8_t* d_pubs = (uint8_t*)ctx->scratch_get("vf_pubs", std::max<size_t>(32, B * (size_t)n_public * 32));
    int32_t* d_valid = (int32_t*)ctx->scratch_get("vf_valid", B * 4);
    if (!d_proofs || !d_pubs || !d_valid) return ctx->fail(NZCB_E_NOMEM, "verify: cannot allocate the device buffers");
    NZ_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
    NZ_CUDA(ctx, cudaMemcpyAsync(d_proofs, proofs, B * sizeof(nzcb_proof), cudaMemcpyHostToDevice, ctx->stream));
    if (n_public)
        NZ_CUDA(ctx, cudaMemcpyAsync(d_pubs, public_le, B * (size_t)n_public * 32, cudaMemcpyHostToDevice, ctx->stream));
    const char* e_serial = getenv("NZCB_VERIFY_SERIAL");
    const char* e_group = getenv("NZCB_VERIFY_GROUP");  // proofs per warp: 1, 2, 4, 8, 16 (default: from the batch size)
    const int serial = e_serial && e_serial[0] == '1';
    // two warps per SM run at nearly the single-proof latency (a third already slows all of them: 1,024 warps of
    // one proof took 86 ms, 256 warps of sixteen 93 ms); beyond that, pack more proofs into a warp
    uint32_t G = 1;
    const size_t slots = (size_t)ctx->sm_count * 2;
    while (G < VERIFY_GMAX && B > slots * G) G *= 2;
    if (e_group) {
        const int g = atoi(e_group);
        if (g == 1 || g == 2 || g == 4 || g == 8 || g == 16) G = (uint32_t)g;
    }
    if (serial) G = 1;
    static const bool attr_set = [] {
        return cudaFuncSetAttribute(k_plonk_verify, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)(VERIFY_GMAX * sizeof(VerifySmem))) == cudaSuccess;
    }();
    if (!attr_set) return ctx->fail(NZCB_E_CUDA, "verify: cannot reserve %zu bytes of shared memory", VERIFY_GMAX * sizeof(VerifySmem));
    {
        const size_t n_groups = (B + G - 1) / G;
        const uint32_t grid = (uint32_t)std::min<size_t>(n_groups, (size_t)ctx->sm_count * 32);
        NZ_LAUNCH(ctx, k_plonk_verify, grid, serial ? 1 : 32, G * sizeof(VerifySmem), vk->d, d_proofs, d_pubs, n_public,
                  (uint32_t)B, d_valid, serial, G);
    }
    NZ_CUDA(ctx, cudaMemcpyAsync(valid, d_valid, B * 4, cudaMemcpyDeviceToHost, ctx->stream));
    NZ_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
    NZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    cudaEventElapsedTime(&ctx->last_ms, ctx->ev0, ctx->ev1);
    return 0;
}

extern "C" int32_t nzcb_pairing_eq(nzcb_ctx* ctx, const uint8_t* g1_affine_lem, const uint8_t* g2_affine_lem, uint32_t n,
                                   int32_t* result, uint8_t* gt_out_lem) {
    if (!ctx || !result || (n && (!g1_affine_lem || !g2_affine_lem))) return NZCB_E_INVALID;
    if (n > 32) return ctx->fail(NZCB_E_INVALID, "pairingEq: at most 32 pairs");
    NZ_CUDA(ctx, cudaSetDevice(ctx->device));
    uint8_t* d = (uint8_t*)ctx->scratch_get("pe_buf", 32 * (64 + 128) + sizeof(Fq12) + 64);
    if (!d) return ctx->fail(NZCB_E_NOMEM, "pairingEq: cannot allocate");
    G1Affine* dP = (G1Affine*)d;
    G2Affine* dQ = (G2Affine*)(d + 32 * 64);
    Fq12* dG = (Fq12*)(d + 32 * 192);
    int32_t* dR = (int32_t*)(d + 32 * 192 + sizeof(Fq12));
    if (n) {
        NZ_CUDA(ctx, cudaMemcpyAsync(dP, g1_affine_lem, (size_t)n * 64, cudaMemcpyHostToDevice, ctx->stream));
        NZ_CUDA(ctx, cudaMemcpyAsync(dQ, g2_affine_lem, (size_t)n * 128, cudaMemcpyHostToDevice, ctx->stream));
    }
    NZ_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
    NZ_LAUNCH(ctx, k_pairing_eq, 1, 32, 0, dP, dQ, n, dR, dG);
    NZ_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
    NZ_CUDA(ctx, cudaMemcpyAsync(result, dR, 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (gt_out_lem) NZ_CUDA(ctx, cudaMemcpyAsync(gt_out_lem, dG, sizeof(Fq12), cudaMemcpyDeviceToHost, ctx->stream));
    NZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    cudaEventElapsedTime(&ctx->last_ms, ctx->ev0, ctx->ev1);
    if (*result < 0) {
        *result = 0;
        return ctx->fail(NZCB_E_INVALID, "pairingEq: a point is not on its curve");
    }
    return 0;
}

extern "C" int32_t nzcb_srs_g2(nzcb_ctx* ctx, const uint8_t tau_le[32], uint8_t out_affine_lem[128]) {
    if (!ctx || !tau_le || !out_affine_lem) return NZCB_E_INVALID;
    NZ_CUDA(ctx, cudaSetDevice(ctx->device));
    Fr k;
    memcpy(k.v, tau_le, 32);
    const Fr m = Fr::modulus();
    bool lt = false;
    for (int i = 7; i >= 0; i--)
        if (k.v[i] != m.v[i]) {
            lt = k.v[i] < m.v[i];
            break;
        }
    if (!lt) return ctx->fail(NZCB_E_INVALID, "srs: tau is not a canonical Fr element");
    G2Affine* d = (G2Affine*)ctx->scratch_get("g2_out", sizeof(G2Affine));
    if (!d) return ctx->fail(NZCB_E_NOMEM, "srs: cannot allocate");
    NZ_LAUNCH(ctx, k_g2_mul, 1, 1, 0, g2_generator(), k, d);
    NZ_CUDA(ctx, cudaMemcpyAsync(out_affine_lem, d, 128, cudaMemcpyDeviceToHost, ctx->stream));
    NZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

// `snarkjs zkey export soliditycalldata` for a PLONK proof: 0x<proof bytes>,["0x<pub_0>",...] (the proof bytes are
// the nzcb_proof layout: nine uncompressed G1 points, seven evaluations, big-endian)
extern "C" int32_t nzcb_proof_to_calldata(const nzcb_proof* p, const uint8_t* public_le, uint32_t n_public, char* buf,
                                          size_t* len) {
    if (!p || !len || (n_public && !public_le)) return NZCB_E_INVALID;
    static const char* hx = "0123456789abcdef";
    std::string o = "0x";
    const uint8_t* raw = (const uint8_t*)p;
    for (size_t i = 0; i < sizeof(nzcb_proof); i++) {
        o += hx[raw[i] >> 4];
        o += hx[raw[i] & 15];
    }
    o += ",[";
    for (uint32_t i = 0; i < n_public; i++) {
        if (i) o += ",";
        o += "\"0x";
        for (int k = 31; k >= 0; k--) {
            const uint8_t b = public_le[(size_t)i * 32 + k];
            o += hx[b >> 4];
            o += hx[b & 15];
        }
        o += "\"";
    }
    o += "]";
    const size_t need = o.size() + 1;
    if (!buf) {
        *len = need;
        return 0;
    }
    if (*len < need) {
        *len = need;
        return NZCB_E_INVALID;
    }
    memcpy(buf, o.c_str(), need);
    *len = need;
    return 0;
}
