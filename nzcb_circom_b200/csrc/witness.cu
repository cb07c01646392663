// Batched witness calculation -- replaces the circom-generated WASM +
// circom_runtime 0.1.17 WitnessCalculator (un-vendored,
// /root/reference/yarn.lock:2496; call sites /root/reference/test/nzcp.js:42 and
// the 72 other calculateWitness calls of test/cbor.js, test/quinSelector.js,
// test/nzcp.js).  The circuit arrives as a witness program (.wprog, emitted by
// nzcb_circom_b200/circom/builder.py): one instruction per wire, sorted by
// dependency level.
//
// Execution model: one CTA per pass; a level's instructions are spread over the
// CTA's threads, levels are separated by __syncthreads.  Wires live in HBM in
// canonical (non-Montgomery) little-endian form, pass-major, so a pass's first
// nWitness wires ARE the payload of its .wtns file and feed the prover without a
// copy.  Constants are Montgomery, so const * wire is a single multiply;
// wire * wire takes two.  IsZero's inverse hint hits a 2 x 1024-entry table for
// the |x| <= 1024 operands the CBOR selectors produce and
// falls back to a Fermat inverse otherwise.
#include "common.cuh"

using namespace nzcb;

enum { OP_LIN = 1, OP_MUL = 2, OP_BITS = 3, OP_INV = 4, OP_ASSERT = 5 };
constexpr uint32_t NZ_INV_TAB = 1024;

struct nzcb_circuit {
    nzcb_ctx* ctx = nullptr;
    uint32_t n_total = 0, n_witness = 0, n_out = 0, n_in = 0, n_consts = 0, n_instr = 0, n_levels = 0, n_code = 0;
    Fr* d_consts = nullptr;      // Montgomery
    uint32_t* d_ioff = nullptr;
    uint32_t* d_lstart = nullptr;
    uint32_t* d_code = nullptr;
    Fr* d_invtab = nullptr;      // canonical 1/k, k = 0..NZ_INV_TAB (entry 0 unused)
};

namespace {

struct ProgView {
    const Fr* consts;
    const uint32_t *ioff, *lstart, *code;
    const Fr* invtab;
    uint32_t n_total, n_out, n_in, n_levels;
};

// canonical value of  k + sum coef_i * w_i ;  advances p past the encoded LC
__device__ __forceinline__ Fr eval_lc(const ProgView& pv, const Fr* __restrict__ W, uint32_t& p) {
    const uint32_t n = pv.code[p], ci = pv.code[p + 1];
    p += 2;
    Fr acc = ci != 0xffffffffu ? pv.consts[ci].from_mont() : Fr::zero();
    for (uint32_t t = 0; t < n; t++) {
        const uint32_t w = pv.code[p], c = pv.code[p + 1];
        p += 2;
        const Fr v = W[w];
        if (c == 0) acc = acc + v;            // coefficient 1
        else if (c == 1) acc = acc - v;       // coefficient -1
        else if (!v.is_zero()) acc = acc + pv.consts[c] * v;  // Montgomery const x canonical wire = canonical
    }
    return acc;
}

__device__ __forceinline__ Fr inv_or_zero(const ProgView& pv, const Fr& v) {
    if (v.is_zero()) return v;
    uint32_t hi = 0;
#pragma unroll
    for (int i = 1; i < 8; i++) hi |= v.v[i];
    if (hi == 0 && v.v[0] <= NZ_INV_TAB) return pv.invtab[v.v[0]];
    const Fr m = v.neg();
    hi = 0;
#pragma unroll
    for (int i = 1; i < 8; i++) hi |= m.v[i];
    if (hi == 0 && m.v[0] <= NZ_INV_TAB) return pv.invtab[m.v[0]].neg();  // 1/(-k) = -(1/k)
    return v.to_mont().inv().from_mont();
}

__global__ void __launch_bounds__(256) k_witness(ProgView pv, const Fr* __restrict__ inputs, Fr* __restrict__ wires,
                                                 int32_t* __restrict__ status, uint32_t B) {
    for (uint32_t pass = blockIdx.x; pass < B; pass += gridDim.x) {
        Fr* W = wires + (size_t)pass * pv.n_total;
        const Fr* in = inputs + (size_t)pass * pv.n_in;
        for (uint32_t i = threadIdx.x; i < pv.n_in; i += blockDim.x) W[1 + pv.n_out + i] = in[i];
        if (threadIdx.x == 0) {
            Fr one = Fr::zero();
            one.v[0] = 1;
            W[0] = one;
        }
        __syncthreads();
        bool failed = false;
        for (uint32_t l = 0; l < pv.n_levels; l++) {
            const uint32_t lo = pv.lstart[l], hi = pv.lstart[l + 1];
            for (uint32_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
                uint32_t p = pv.ioff[i];
                const uint32_t op = pv.code[p];
                if (op == OP_LIN) {
                    const uint32_t dst = pv.code[p + 1];
                    p += 2;
                    W[dst] = eval_lc(pv, W, p);
                } else if (op == OP_MUL) {
                    const uint32_t dst = pv.code[p + 1];
                    p += 2;
                    const Fr a = eval_lc(pv, W, p);
                    const Fr b = eval_lc(pv, W, p);
                    const Fr c = eval_lc(pv, W, p);
                    Fr ab = Fr::zero();
                    if (!a.is_zero() && !b.is_zero()) ab = (a * b) * Fr::r2();
                    W[dst] = ab + c;
                } else if (op == OP_BITS) {
                    const uint32_t dst = pv.code[p + 1], src = pv.code[p + 2], n = pv.code[p + 3];
                    const Fr v = W[src];
                    for (uint32_t k = 0; k < n; k++) {
                        Fr bit = Fr::zero();
                        bit.v[0] = k < 256 ? (v.v[k >> 5] >> (k & 31)) & 1u : 0u;
                        W[dst + k] = bit;
                    }
                } else if (op == OP_INV) {
                    W[pv.code[p + 1]] = inv_or_zero(pv, W[pv.code[p + 2]]);
                } else {  // OP_ASSERT
                    p += 1;
                    const Fr a = eval_lc(pv, W, p);
                    const Fr b = eval_lc(pv, W, p);
                    const Fr c = eval_lc(pv, W, p);
                    Fr ab = Fr::zero();
                    if (!a.is_zero() && !b.is_zero()) ab = (a * b) * Fr::r2();
                    if (ab != c) failed = true;
                }
            }
            __syncthreads();
        }
        if (failed) atomicExch(&status[pass], NZCB_E_ASSERT);
        __syncthreads();
    }
}

__global__ void k_invtab(Fr* tab) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k > NZ_INV_TAB) return;
    tab[k] = k == 0 ? Fr::zero() : Fr::from_u64(k).inv().from_mont();
}

__global__ void k_consts_to_mont(Fr* c, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) c[i] = c[i].to_mont();
}

}  // namespace

extern "C" void nzcb_circuit_free(nzcb_circuit* c) {
    if (!c) return;
    if (c->ctx) {
        cudaSetDevice(c->ctx->device);
        cudaStreamSynchronize(c->ctx->stream);
    }
    cudaFree(c->d_consts);
    cudaFree(c->d_ioff);
    cudaFree(c->d_lstart);
    cudaFree(c->d_code);
    cudaFree(c->d_invtab);
    delete c;
}

extern "C" int32_t nzcb_circuit_info(const nzcb_circuit* c, uint32_t* n_witness, uint32_t* n_inputs, uint32_t* n_outputs) {
    if (!c) return NZCB_E_INVALID;
    if (n_witness) *n_witness = c->n_witness;
    if (n_inputs) *n_inputs = c->n_in;
    if (n_outputs) *n_outputs = c->n_out;
    return 0;
}

#define WC_CUDA(call)                                                                                      \
    do {                                                                                                   \
        cudaError_t e__ = (call);                                                                          \
        if (e__ != cudaSuccess) {                                                                          \
            ctx->fail(NZCB_E_CUDA, "CUDA error %s at %s:%d", cudaGetErrorString(e__), __FILE__, __LINE__); \
            nzcb_circuit_free(c);                                                                          \
            return NZCB_E_CUDA;                                                                            \
        }                                                                                                  \
    } while (0)

extern "C" int32_t nzcb_circuit_load(nzcb_ctx* ctx, const uint8_t* data, size_t len, nzcb_circuit** out) {
    if (!ctx || !data || !out) return NZCB_E_INVALID;
    *out = nullptr;
    if (len < 40 || memcmp(data, "NZWP", 4) != 0) return ctx->fail(NZCB_E_INVALID, "witness program: bad magic");
    uint32_t h[9];
    memcpy(h, data + 4, 36);
    if (h[0] != 1) return ctx->fail(NZCB_E_INVALID, "witness program: unsupported version %u", h[0]);
    nzcb_circuit* c = new nzcb_circuit();
    c->ctx = ctx;
    c->n_total = h[1]; c->n_witness = h[2]; c->n_out = h[3]; c->n_in = h[4];
    c->n_consts = h[5]; c->n_instr = h[6]; c->n_levels = h[7]; c->n_code = h[8];
    const size_t need = 40 + (size_t)c->n_consts * 32 + (size_t)c->n_instr * 4 + ((size_t)c->n_levels + 1) * 4 +
                        (size_t)c->n_code * 4;
    if (need != len || c->n_witness > c->n_total || 1 + (uint64_t)c->n_out + c->n_in > c->n_witness || c->n_consts < 2) {
        delete c;
        return ctx->fail(NZCB_E_INVALID, "witness program: inconsistent header");
    }
    const uint8_t* p_consts = data + 40;
    const uint8_t* p_ioff = p_consts + (size_t)c->n_consts * 32;
    const uint8_t* p_lstart = p_ioff + (size_t)c->n_instr * 4;
    const uint8_t* p_code = p_lstart + ((size_t)c->n_levels + 1) * 4;
    // validate once on the host so the kernel can trust every index
    {
        std::vector<uint32_t> code(c->n_code), ioff(c->n_instr), ls(c->n_levels + 1);
        memcpy(code.data(), p_code, (size_t)c->n_code * 4);
        memcpy(ioff.data(), p_ioff, (size_t)c->n_instr * 4);
        memcpy(ls.data(), p_lstart, ((size_t)c->n_levels + 1) * 4);
        bool ok = ls[0] == 0 && ls[c->n_levels] == c->n_instr;
        for (uint32_t l = 0; ok && l < c->n_levels; l++) ok = ls[l] <= ls[l + 1];
        auto lc_ok = [&](uint32_t& p) {
            if (p + 2 > c->n_code) return false;
            const uint32_t n = code[p], ci = code[p + 1];
            if (ci != 0xffffffffu && ci >= c->n_consts) return false;
            p += 2;
            if ((uint64_t)p + 2ull * n > c->n_code) return false;
            for (uint32_t t = 0; t < n; t++, p += 2)
                if (code[p] >= c->n_total || code[p + 1] >= c->n_consts) return false;
            return true;
        };
        for (uint32_t i = 0; ok && i < c->n_instr; i++) {
            uint32_t p = ioff[i];
            if (p + 1 > c->n_code) { ok = false; break; }
            const uint32_t op = code[p];
            if (op == OP_LIN) {
                ok = p + 2 <= c->n_code && code[p + 1] < c->n_total; p += 2; ok = ok && lc_ok(p);
            } else if (op == OP_MUL) {
                ok = p + 2 <= c->n_code && code[p + 1] < c->n_total; p += 2; ok = ok && lc_ok(p) && lc_ok(p) && lc_ok(p);
            } else if (op == OP_BITS) {
                ok = p + 4 <= c->n_code && code[p + 2] < c->n_total && (uint64_t)code[p + 1] + code[p + 3] <= c->n_total;
            } else if (op == OP_INV) {
                ok = p + 3 <= c->n_code && code[p + 1] < c->n_total && code[p + 2] < c->n_total;
            } else if (op == OP_ASSERT) {
                p += 1; ok = lc_ok(p) && lc_ok(p) && lc_ok(p);
            } else ok = false;
        }
        if (!ok) {
            delete c;
            return ctx->fail(NZCB_E_INVALID, "witness program: malformed instruction stream");
        }
    }
    WC_CUDA(cudaSetDevice(ctx->device));
    WC_CUDA(cudaMalloc(&c->d_consts, (size_t)c->n_consts * 32));
    WC_CUDA(cudaMalloc(&c->d_ioff, std::max<size_t>(4, (size_t)c->n_instr * 4)));
    WC_CUDA(cudaMalloc(&c->d_lstart, ((size_t)c->n_levels + 1) * 4));
    WC_CUDA(cudaMalloc(&c->d_code, std::max<size_t>(4, (size_t)c->n_code * 4)));
    WC_CUDA(cudaMalloc(&c->d_invtab, (NZ_INV_TAB + 1) * sizeof(Fr)));
    WC_CUDA(cudaMemcpyAsync(c->d_consts, p_consts, (size_t)c->n_consts * 32, cudaMemcpyHostToDevice, ctx->stream));
    WC_CUDA(cudaMemcpyAsync(c->d_ioff, p_ioff, (size_t)c->n_instr * 4, cudaMemcpyHostToDevice, ctx->stream));
    WC_CUDA(cudaMemcpyAsync(c->d_lstart, p_lstart, ((size_t)c->n_levels + 1) * 4, cudaMemcpyHostToDevice, ctx->stream));
    WC_CUDA(cudaMemcpyAsync(c->d_code, p_code, (size_t)c->n_code * 4, cudaMemcpyHostToDevice, ctx->stream));
    k_consts_to_mont<<<div_up(c->n_consts, 256), 256, 0, ctx->stream>>>(c->d_consts, c->n_consts);
    k_invtab<<<div_up(NZ_INV_TAB + 1, 128), 128, 0, ctx->stream>>>(c->d_invtab);
    ctx->launches += 2;
    WC_CUDA(cudaGetLastError());
    WC_CUDA(cudaStreamSynchronize(ctx->stream));
    *out = c;
    return 0;
}

namespace nzcb {
// runs B passes; wires for pass i start at *d_wires + i * n_total (canonical LE).  Asynchronous on ctx->stream.
int witness_dev(nzcb_ctx* ctx, const nzcb_circuit* c, const Fr* d_inputs, size_t B, Fr* d_wires, int32_t* d_status) {
    ProgView pv;
    pv.consts = c->d_consts; pv.ioff = c->d_ioff; pv.lstart = c->d_lstart; pv.code = c->d_code; pv.invtab = c->d_invtab;
    pv.n_total = c->n_total; pv.n_out = c->n_out; pv.n_in = c->n_in; pv.n_levels = c->n_levels;
    NZ_CUDA(ctx, cudaMemsetAsync(d_status, 0, B * sizeof(int32_t), ctx->stream));
    const uint32_t grid = (uint32_t)std::min<size_t>(B, (size_t)ctx->sm_count * 8);
    NZ_LAUNCH(ctx, k_witness, grid, 256, 0, pv, d_inputs, d_wires, d_status, (uint32_t)B);
    return 0;
}
uint32_t circuit_n_total(const nzcb_circuit* c) { return c->n_total; }
uint32_t circuit_n_witness(const nzcb_circuit* c) { return c->n_witness; }
uint32_t circuit_n_in(const nzcb_circuit* c) { return c->n_in; }
}  // namespace nzcb

extern "C" int32_t nzcb_witness_batch(nzcb_ctx* ctx, const nzcb_circuit* c, const uint8_t* inputs_le, size_t B,
                                      uint8_t* wtns_out, int32_t* status) {
    if (!ctx || !c || (!inputs_le && c->n_in) || !status) return NZCB_E_INVALID;
    if (c->ctx != ctx) return ctx->fail(NZCB_E_INVALID, "circuit was loaded on a different context");
    if (B == 0) return 0;
    NZ_CUDA(ctx, cudaSetDevice(ctx->device));
    // bound the device footprint: process the batch in chunks of at most ~8 GiB of wires
    const size_t per_pass = (size_t)c->n_total * sizeof(Fr);
    size_t chunk = std::max<size_t>(1, ((size_t)8 << 30) / per_pass);
    if (chunk > B) chunk = B;
    Fr* d_w = (Fr*)ctx->scratch_get("wt_wires", chunk * per_pass);
    Fr* d_in = (Fr*)ctx->scratch_get("wt_inputs", std::max<size_t>(32, chunk * (size_t)c->n_in * sizeof(Fr)));
    int32_t* d_st = (int32_t*)ctx->scratch_get("wt_status", chunk * sizeof(int32_t));
    if (!d_w || !d_in || !d_st) return ctx->fail(NZCB_E_NOMEM, "witness: cannot allocate %zu device bytes", chunk * per_pass);
    NZ_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
    for (size_t done = 0; done < B; done += chunk) {
        const size_t nb = std::min(chunk, B - done);
        if (c->n_in)
            NZ_CUDA(ctx, cudaMemcpyAsync(d_in, inputs_le + done * (size_t)c->n_in * 32, nb * (size_t)c->n_in * 32,
                                         cudaMemcpyHostToDevice, ctx->stream));
        NZ_TRY(witness_dev(ctx, c, d_in, nb, d_w, d_st));
        NZ_CUDA(ctx, cudaMemcpyAsync(status + done, d_st, nb * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
        if (wtns_out) {
            if (c->n_total == c->n_witness) {
                NZ_CUDA(ctx, cudaMemcpyAsync(wtns_out + done * (size_t)c->n_witness * 32, d_w, nb * per_pass,
                                             cudaMemcpyDeviceToHost, ctx->stream));
            } else {
                NZ_CUDA(ctx, cudaMemcpy2DAsync(wtns_out + done * (size_t)c->n_witness * 32, (size_t)c->n_witness * 32, d_w,
                                               per_pass, (size_t)c->n_witness * 32, nb, cudaMemcpyDeviceToHost,
                                               ctx->stream));
            }
        }
        NZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    NZ_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
    NZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    cudaEventElapsedTime(&ctx->last_ms, ctx->ev0, ctx->ev1);
    return 0;
}
