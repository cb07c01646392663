// Context lifetime and small device-memory helpers of the C ABI (include/nzcb.h).
#include "common.cuh"

using namespace nzcb;

static thread_local char g_noctx_err[256] = "nzcb: no context";

extern "C" int32_t nzcb_ctx_create(int32_t device_id, nzcb_ctx** out) {
    if (!out) return NZCB_E_INVALID;
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        snprintf(g_noctx_err, sizeof(g_noctx_err),
                 "nzcb: no CUDA device available (%s); there is no CPU fallback",
                 e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
        cudaGetLastError();
        return NZCB_E_CUDA;
    }
    if (device_id < 0 || device_id >= ndev) {
        snprintf(g_noctx_err, sizeof(g_noctx_err), "nzcb: device %d out of range (have %d)", device_id, ndev);
        return NZCB_E_INVALID;
    }
    cudaDeviceProp prop;
    if (cudaSetDevice(device_id) != cudaSuccess || cudaGetDeviceProperties(&prop, device_id) != cudaSuccess) {
        snprintf(g_noctx_err, sizeof(g_noctx_err), "nzcb: cannot select device %d", device_id);
        cudaGetLastError();
        return NZCB_E_CUDA;
    }
    if (prop.major < 10) {
        snprintf(g_noctx_err, sizeof(g_noctx_err), "nzcb: device %d is sm_%d%d; this library is built for sm_100a only",
                 device_id, prop.major, prop.minor);
        return NZCB_E_CUDA;
    }
    nzcb_ctx* ctx = new nzcb_ctx();
    ctx->device = device_id;
    ctx->sm_count = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreate(&ctx->ev0) != cudaSuccess || cudaEventCreate(&ctx->ev1) != cudaSuccess) {
        snprintf(g_noctx_err, sizeof(g_noctx_err), "nzcb: cannot create stream/events on device %d", device_id);
        delete ctx;
        cudaGetLastError();
        return NZCB_E_CUDA;
    }
    *out = ctx;
    return 0;
}

namespace nzcb {
nzcb_ctx* ctx_lane(nzcb_ctx* root, int i) {
    if (i < 0 || i >= 16) return nullptr;
    std::lock_guard<std::mutex> g(root->mu);
    while ((int)root->lanes.size() <= i) {
        nzcb_ctx* l = new nzcb_ctx();
        l->parent = root;
        l->device = root->device;
        l->sm_count = root->sm_count;
        cudaSetDevice(root->device);
        if (cudaStreamCreateWithFlags(&l->stream, cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreate(&l->ev0) != cudaSuccess || cudaEventCreate(&l->ev1) != cudaSuccess) {
            cudaGetLastError();
            if (l->stream) cudaStreamDestroy(l->stream);
            if (l->ev0) cudaEventDestroy(l->ev0);
            if (l->ev1) cudaEventDestroy(l->ev1);
            delete l;
            return nullptr;
        }
        root->lanes.push_back(l);
    }
    return root->lanes[i];
}
}  // namespace nzcb

extern "C" void nzcb_ctx_free(nzcb_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    for (nzcb_ctx* l : ctx->lanes) nzcb_ctx_free(l);
    ctx->lanes.clear();
    nzcb::msm_split_release(ctx);
    for (int b = 0; b < 2; b++) {
        if (ctx->wt_pin[b]) cudaFreeHost(ctx->wt_pin[b]);
        if (ctx->wt_ev[b]) cudaEventDestroy(ctx->wt_ev[b]);
    }
    if (ctx->wt_copy) cudaStreamDestroy(ctx->wt_copy);
    for (auto& kv : ctx->twiddles) cudaFree(kv.second);
    for (auto& kv : ctx->scratch)
        if (kv.second.first) cudaFree(kv.second.first);
    for (auto& e : ctx->prof_ev) {
        cudaEventDestroy(e.first);
        cudaEventDestroy(e.second);
    }
    if (ctx->side) cudaStreamDestroy(ctx->side);
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

extern "C" int32_t nzcb_profile(nzcb_ctx* ctx, int32_t enable) {
    if (!ctx) return NZCB_E_INVALID;
    NZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->prof_on = enable != 0;
    ctx->prof_used = 0;
    ctx->prof_modmul = 0;
    for (nzcb_ctx* l : ctx->lanes) nzcb_profile(l, enable);
    return 0;
}
// bucket additions actually executed by the timed launches since the last read (call before nzcb_profile_read)
extern "C" int32_t nzcb_profile_entries(nzcb_ctx* ctx, double* additions) {
    if (!ctx || !additions) return NZCB_E_INVALID;
    NZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    double n = 0;
    for (size_t i = 0; i < ctx->prof_used && i < ctx->prof_entries.size(); i++) n += ctx->prof_entries[i];
    for (nzcb_ctx* l : ctx->lanes) {
        double ln = 0;
        nzcb_profile_entries(l, &ln);
        n += ln;
    }
    *additions = n;
    return 0;
}

extern "C" int32_t nzcb_profile_read(nzcb_ctx* ctx, uint64_t* launches, double* total_ms, double* alg_modmul) {
    if (!ctx) return NZCB_E_INVALID;
    NZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    double ms = 0;
    for (size_t i = 0; i < ctx->prof_used; i++) {
        float t = 0;
        cudaEventElapsedTime(&t, ctx->prof_ev[i].first, ctx->prof_ev[i].second);
        ms += t;
    }
    uint64_t n = ctx->prof_used;
    double mm = ctx->prof_modmul;
    ctx->prof_used = 0;
    ctx->prof_modmul = 0;
    for (nzcb_ctx* l : ctx->lanes) {
        uint64_t ln = 0;
        double lms = 0, lmm = 0;
        nzcb_profile_read(l, &ln, &lms, &lmm);
        n += ln;
        ms += lms;
        mm += lmm;
    }
    if (launches) *launches = n;
    if (total_ms) *total_ms = ms;
    if (alg_modmul) *alg_modmul = mm;
    return 0;
}

extern "C" const char* nzcb_last_error(const nzcb_ctx* ctx) { return ctx ? ctx->err : g_noctx_err; }
extern "C" uint64_t nzcb_launch_count(const nzcb_ctx* ctx) {
    if (!ctx) return 0;
    uint64_t n = ctx->launches;
    for (const nzcb_ctx* l : ctx->lanes) n += l->launches;
    return n;
}
extern "C" float nzcb_last_device_ms(const nzcb_ctx* ctx) { return ctx ? ctx->last_ms : 0.f; }

extern "C" int32_t nzcb_dev_alloc(nzcb_ctx* ctx, size_t bytes, void** dptr) {
    if (!ctx || !dptr) return NZCB_E_INVALID;
    NZ_CUDA(ctx, cudaSetDevice(ctx->device));
    if (cudaMalloc(dptr, bytes ? bytes : 1) != cudaSuccess) {
        cudaGetLastError();
        return ctx->fail(NZCB_E_NOMEM, "cannot allocate %zu device bytes", bytes);
    }
    return 0;
}
extern "C" int32_t nzcb_dev_free(nzcb_ctx* ctx, void* dptr) {
    if (!ctx) return NZCB_E_INVALID;
    NZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    NZ_CUDA(ctx, cudaFree(dptr));
    return 0;
}
extern "C" int32_t nzcb_dev_upload(nzcb_ctx* ctx, void* dptr, const void* host, size_t bytes) {
    if (!ctx || !dptr || !host) return NZCB_E_INVALID;
    NZ_CUDA(ctx, cudaMemcpyAsync(dptr, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    NZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}
extern "C" int32_t nzcb_dev_download(nzcb_ctx* ctx, void* host, const void* dptr, size_t bytes) {
    if (!ctx || !dptr || !host) return NZCB_E_INVALID;
    NZ_CUDA(ctx, cudaMemcpyAsync(host, dptr, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    NZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}
