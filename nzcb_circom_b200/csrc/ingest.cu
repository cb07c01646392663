// Pass ingest on the device: "NZCP:/1/<base32>" -> COSE_Sign1 -> ToBeSigned -> the circuit's main inputs.
//
// Replaces, for a batch of passes, what the reference's tests do on the host before calculateWitness:
//   /root/reference/test/helpers/nzcp.js:9-24     base32ToBytes
//   /root/reference/test/helpers/nzcp.js:58-105   decodeCBORStream  (only as far as decodeCOSE looks)
//   /root/reference/test/helpers/nzcp.js:141-172  decodeBytes / decodeCOSE / getCOSE
//   /root/reference/test/helpers/nzcp.js:123-137,180-206  encodeBytes / encodeToBeSigned
//   /root/reference/test/helpers/utils.js:2,49,71,87      bufferToBitArray, fitBytes, evmRearrangeBytes
//   /root/reference/test/nzcp.js:36-41            { toBeSigned, toBeSignedLen, data }
// One CTA per pass: the threads decode base32 into shared memory, one thread walks the four COSE fields (a few
// header bytes), then all threads emit ToBeSigned and the marshalled inputs (one 16-byte store per thread and
// half element, coalesced).  HBM: ~0.6 KB read and (8 maxLen + 161) x 32 B written per pass.
#include "common.cuh"
#include "ingest.cuh"

using namespace nzcb;

namespace {
constexpr int ING_THREADS = 128;

__global__ void __launch_bounds__(ING_THREADS)
k_pass_ingest(const uint8_t* __restrict__ uris, const uint32_t* __restrict__ uri_off, const uint8_t* __restrict__ data20,
              uint32_t max_len, uint32_t B, uint8_t* __restrict__ tbs_out, uint32_t* __restrict__ tbs_len_out,
              uint4* __restrict__ inputs, int32_t* __restrict__ status) {
    __shared__ uint8_t raw[ING_MAX_RAW + 3];
    __shared__ uint8_t tbs[ING_MAX_TBS];
    __shared__ uint8_t dat[20];
    __shared__ CoseFields f;
    const uint32_t n_in = 8 * max_len + 161;
    for (uint32_t b = blockIdx.x; b < B; b += gridDim.x) {
        const uint32_t u0 = uri_off[b], total = uri_off[b + 1] - u0;
        const uint32_t n = total > 8 ? total - 8 : 0;  // passURI.substring(8), prefix unchecked
        const uint8_t* sym = uris + u0 + 8;
        const bool too_long = n > ING_MAX_CHARS;
        const uint32_t n_raw = too_long ? 0 : (n * 5 + 7) / 8;
        int bad = too_long;
        if (!too_long)
            for (uint32_t i = threadIdx.x; i < n; i += ING_THREADS) bad |= b32_val(sym[i]) < 0;
        bad = __syncthreads_or(bad);
        if (!bad)
            for (uint32_t j = threadIdx.x; j < n_raw; j += ING_THREADS) raw[j] = b32_out_byte(sym, n, j);
        if (threadIdx.x < 20) dat[threadIdx.x] = data20 ? data20[(size_t)b * 20 + threadIdx.x] : 0;
        __syncthreads();
        if (threadIdx.x == 0) {
            CoseFields g = {};
            if (!bad) g = parse_cose(raw, n_raw);
            f = g;
        }
        __syncthreads();
        const CoseFields g = f;
        for (uint32_t t = threadIdx.x; t < max_len; t += ING_THREADS) {
            const uint8_t o = tbs_byte_at(g, raw, t);
            tbs[t] = o;
            if (tbs_out) tbs_out[(size_t)b * max_len + t] = o;
        }
        if (threadIdx.x == 0) {
            if (tbs_len_out) tbs_len_out[b] = g.ok ? g.tbs_len : 0;
            status[b] = g.ok ? 0 : NZCB_E_INVALID;
        }
        __syncthreads();
        if (inputs) {
            // one 16-byte store per thread and half element: a warp writes 512 contiguous bytes
            uint4* dst = inputs + (size_t)b * n_in * 2;
            for (uint32_t h = threadIdx.x; h < 2 * n_in; h += ING_THREADS)
                dst[h] = make_uint4((h & 1) ? 0u : ingest_input_value(g, tbs, dat, max_len, h >> 1), 0, 0, 0);
        }
        __syncthreads();
    }
}

int ingest_dev(nzcb_ctx* ctx, const uint8_t* d_uris, const uint32_t* d_off, const uint8_t* d_data, uint32_t max_len,
               size_t B, uint8_t* d_tbs, uint32_t* d_len, uint4* d_inputs, int32_t* d_status) {
    const uint32_t grid = (uint32_t)std::min<size_t>(B, (size_t)ctx->sm_count * 16);
    NZ_LAUNCH(ctx, k_pass_ingest, grid, ING_THREADS, 0, d_uris, d_off, d_data, max_len, (uint32_t)B, d_tbs, d_len,
              d_inputs, d_status);
    return 0;
}

int ingest_impl(nzcb_ctx* ctx, const uint8_t* uris, const uint32_t* uri_off, size_t B, const uint8_t* data20,
                uint32_t max_len, uint8_t* tbs_out, uint32_t* tbs_len, uint8_t* inputs_host, void* inputs_dev,
                int32_t* status) {
    if (!ctx || !uri_off || !status || (!uris && B && uri_off[B])) return NZCB_E_INVALID;
    if (max_len == 0 || max_len > ING_MAX_TBS) return ctx->fail(NZCB_E_INVALID, "ingest: maxLen %u out of range (1..%u)", max_len, ING_MAX_TBS);
    if (B == 0) return 0;
    for (size_t i = 0; i < B; i++)
        if (uri_off[i + 1] < uri_off[i]) return ctx->fail(NZCB_E_INVALID, "ingest: uri_off must be non-decreasing");
    NZ_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t n_chars = uri_off[B], n_in = 8 * (size_t)max_len + 161;
    uint8_t* d_uris = (uint8_t*)ctx->scratch_get("ing_uris", n_chars + 16);
    uint32_t* d_off = (uint32_t*)ctx->scratch_get("ing_off", (B + 1) * 4);
    uint8_t* d_data = data20 ? (uint8_t*)ctx->scratch_get("ing_data", B * 20) : nullptr;
    uint8_t* d_tbs = tbs_out ? (uint8_t*)ctx->scratch_get("ing_tbs", B * max_len) : nullptr;
    uint32_t* d_len = (uint32_t*)ctx->scratch_get("ing_len", B * 4);
    int32_t* d_st = (int32_t*)ctx->scratch_get("ing_status", B * 4);
    uint4* d_inputs = (uint4*)inputs_dev;
    if (inputs_host) d_inputs = (uint4*)ctx->scratch_get("ing_inputs", B * n_in * 32);
    if (!d_uris || !d_off || !d_len || !d_st || (data20 && !d_data) || (tbs_out && !d_tbs) || (inputs_host && !d_inputs))
        return ctx->fail(NZCB_E_NOMEM, "ingest: cannot allocate the device buffers");
    NZ_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
    if (n_chars) NZ_CUDA(ctx, cudaMemcpyAsync(d_uris, uris, n_chars, cudaMemcpyHostToDevice, ctx->stream));
    NZ_CUDA(ctx, cudaMemcpyAsync(d_off, uri_off, (B + 1) * 4, cudaMemcpyHostToDevice, ctx->stream));
    if (data20) NZ_CUDA(ctx, cudaMemcpyAsync(d_data, data20, B * 20, cudaMemcpyHostToDevice, ctx->stream));
    NZ_TRY(ingest_dev(ctx, d_uris, d_off, d_data, max_len, B, d_tbs, d_len, d_inputs, d_st));
    NZ_CUDA(ctx, cudaMemcpyAsync(status, d_st, B * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (tbs_len) NZ_CUDA(ctx, cudaMemcpyAsync(tbs_len, d_len, B * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (tbs_out) NZ_CUDA(ctx, cudaMemcpyAsync(tbs_out, d_tbs, B * max_len, cudaMemcpyDeviceToHost, ctx->stream));
    if (inputs_host) NZ_CUDA(ctx, cudaMemcpyAsync(inputs_host, d_inputs, B * n_in * 32, cudaMemcpyDeviceToHost, ctx->stream));
    NZ_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
    NZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    cudaEventElapsedTime(&ctx->last_ms, ctx->ev0, ctx->ev1);
    return 0;
}
}  // namespace

extern "C" int32_t nzcb_pass_ingest_batch(nzcb_ctx* ctx, const uint8_t* uris, const uint32_t* uri_off, size_t B,
                                          const uint8_t* data20, uint32_t max_len, uint8_t* tbs_out, uint32_t* tbs_len,
                                          uint8_t* inputs_le, int32_t* status) {
    return ingest_impl(ctx, uris, uri_off, B, data20, max_len, tbs_out, tbs_len, inputs_le, nullptr, status);
}

extern "C" int32_t nzcb_pass_ingest_batch_dev(nzcb_ctx* ctx, const uint8_t* uris, const uint32_t* uri_off, size_t B,
                                              const uint8_t* data20, uint32_t max_len, void* d_inputs_le,
                                              int32_t* status) {
    if (!d_inputs_le) return NZCB_E_INVALID;
    return ingest_impl(ctx, uris, uri_off, B, data20, max_len, nullptr, nullptr, nullptr, d_inputs_le, status);
}

extern "C" int32_t nzcb_plonk_fullprove_uri_batch(nzcb_ctx* ctx, const nzcb_circuit* cir, const nzcb_zkey* zk,
                                                  const uint8_t* uris, const uint32_t* uri_off, size_t B,
                                                  const uint8_t* data20, uint32_t max_len, const uint8_t* blinders_le,
                                                  nzcb_proof* out, uint8_t* public_le, int32_t* status) {
    if (!ctx || !cir || !zk || !out || !status) return NZCB_E_INVALID;
    uint32_t n_w = 0, n_in = 0, n_out = 0;
    NZ_TRY(nzcb_circuit_info(cir, &n_w, &n_in, &n_out));
    if (n_in != 8 * max_len + 161)
        return ctx->fail(NZCB_E_INVALID, "fullProve: the circuit takes %u inputs, a pass of maxLen %u marshals to %u", n_in,
                         max_len, 8 * max_len + 161);
    if (B == 0) return 0;
    NZ_CUDA(ctx, cudaSetDevice(ctx->device));
    void* d_inputs = ctx->scratch_get("ing_fp_inputs", B * (size_t)n_in * 32);
    if (!d_inputs) return ctx->fail(NZCB_E_NOMEM, "fullProve: cannot allocate the input buffer");
    std::vector<int32_t> ing(B);
    NZ_TRY(ingest_impl(ctx, uris, uri_off, B, data20, max_len, nullptr, nullptr, nullptr, d_inputs, ing.data()));
    const float ing_ms = ctx->last_ms;
    const int32_t rc = nzcb_plonk_fullprove_batch_dev(ctx, cir, zk, d_inputs, B, blinders_le, out, public_le, status);
    ctx->last_ms += ing_ms;
    // an undecodable pass reached the circuit with toBeSignedLen = 0xFFFF and was rejected there; report why
    for (size_t i = 0; i < B; i++)
        if (ing[i] != 0) {
            status[i] = NZCB_E_INVALID;
            memset(out + i, 0, sizeof(nzcb_proof));
        }
    return rc;
}
