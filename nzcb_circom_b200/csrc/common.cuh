// Context, error plumbing and device workspace shared by every translation unit
// of libnzcb.so.  One ctx = one GPU = one stream (include/nzcb.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <string.h>
#include <map>
#include <mutex>
#include <string>
#include <vector>
#include "../../include/nzcb.h"
#include "g1.cuh"

namespace nzcb {
struct NttTables;
struct MsmWorkspace;
}  // namespace nzcb

// One ctx = one GPU.  The root ctx owns the twiddle tables; "lanes" are child contexts (own stream, scratch
// arenas, events, error text) that the batch entry points drive from one host thread each, so that the
// latency-bound tails of one proof overlap the throughput-bound kernels of another.
struct nzcb_ctx {
    nzcb_ctx* parent = nullptr;          // non-null for a lane
    std::vector<nzcb_ctx*> lanes;        // root only
    std::mutex mu;                       // root only: guards the twiddle map
    int device = 0;
    int sm_count = 148;
    // latency mode: this ctx commits only its slice of every fixed-base MSM; the partial sums of all
    // ranks are exchanged through `split_allgather` (NCCL / P2P on the caller's side) and added up.  Root only.
    int split_rank = 0, split_world = 1;
    int (*split_allgather)(void* user, const void* send, void* recv, size_t bytes) = nullptr;
    void* split_user = nullptr;
    // the same exchange on the device: an NCCL communicator of our own (libnccl dlopen'ed at run time, msm.cu);
    // the partial sums never leave HBM until the summed commitment is read back
    void* split_nccl_comm = nullptr;
    // large witness batches from host memory: pinned staging pair + copy stream, so that the upload of chunk k + 1
    // (host memcpy into pinned memory, then DMA) overlaps the witness kernel of chunk k.  Root only, grow-only.
    void* wt_pin[2] = {nullptr, nullptr};
    size_t wt_pin_bytes = 0;
    cudaStream_t wt_copy = nullptr;
    cudaEvent_t wt_ev[2] = {nullptr, nullptr};
    cudaStream_t stream = nullptr;
    cudaStream_t side = nullptr;         // second stream of this ctx / lane: commitments overlap the transforms of a round
    cudaEvent_t ev_fork = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    float last_ms = 0.f;
    uint64_t launches = 0;
    char err[512] = {0};
    // optional CUDA-event timing of the dominant kernel (k_msm_accum) for bench.py's roofline
    bool prof_on = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_ev;
    size_t prof_used = 0;
    double prof_modmul = 0;  // algorithmic modmul of the timed launches (160 per MSM point)
    std::vector<uint32_t> prof_entries;  // bucket additions actually executed per timed launch
    // twiddle tables per (log_n, inverse)
    std::map<uint32_t, nzcb::Fr*> twiddles;
    // grow-only scratch arenas keyed by name, so steady-state proving never mallocs
    std::map<std::string, std::pair<void*, size_t>> scratch;

    nzcb_ctx* root() { return parent ? parent : this; }
    int fail(int code, const char* fmt, ...) {
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(err, sizeof(err), fmt, ap);
        va_end(ap);
        return code;
    }
    // named scratch buffer of at least `bytes`; contents NOT preserved on growth
    void* scratch_get(const char* name, size_t bytes) {
        auto& e = scratch[name];
        if (e.second >= bytes && e.first) return e.first;
        if (e.first) {
            cudaStreamSynchronize(stream);
            cudaFree(e.first);
            e.first = nullptr;
            e.second = 0;
        }
        void* p = nullptr;
        size_t want = bytes < 256 ? 256 : bytes;
        if (cudaMalloc(&p, want) != cudaSuccess) {
            cudaGetLastError();
            return nullptr;
        }
        e.first = p;
        e.second = want;
        return p;
    }
};

#define NZ_CUDA(ctx, call)                                                                          \
    do {                                                                                            \
        cudaError_t e__ = (call);                                                                   \
        if (e__ != cudaSuccess)                                                                     \
            return (ctx)->fail(NZCB_E_CUDA, "CUDA error %s at %s:%d", cudaGetErrorString(e__), __FILE__, \
                               __LINE__);                                                           \
    } while (0)

#define NZ_TRY(expr)             \
    do {                         \
        int rc__ = (expr);       \
        if (rc__ != 0) return rc__; \
    } while (0)

// launch helper: counts launches (bench.py gpu_launches) and checks the launch
#define NZ_LAUNCH(ctx, kernel, grid, block, smem, ...)                                              \
    do {                                                                                            \
        kernel<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);                            \
        (ctx)->launches++;                                                                          \
        cudaError_t e__ = cudaGetLastError();                                                       \
        if (e__ != cudaSuccess)                                                                     \
            return (ctx)->fail(NZCB_E_CUDA, "launch of %s failed: %s (%s:%d)", #kernel,             \
                               cudaGetErrorString(e__), __FILE__, __LINE__);                        \
    } while (0)

namespace nzcb {

static inline unsigned div_up(size_t a, size_t b) { return (unsigned)((a + b - 1) / b); }

// --- internal device-level entry points (all asynchronous on ctx->stream) ---
// api.cu : lane `i` of a root ctx (created on first use)
nzcb_ctx* ctx_lane(nzcb_ctx* root, int i);
// ntt.cu
int ntt_dev(nzcb_ctx* ctx, Fr* d_data, uint32_t log_n, bool inverse);
int ntt_dev_tab(nzcb_ctx* ctx, Fr* d_data, uint32_t log_n, bool inverse, const Fr* d_out_factors);
// msm.cu : result left in d_out (one G1XYZZ) ; scalars 8 x u32 each.  One-shot bases (window mode).
int msm_dev(nzcb_ctx* ctx, const G1Affine* d_bases, const uint32_t* d_scalars, size_t n, bool scalars_mont,
            G1XYZZ* d_out);
// fixed bases with the window shifts precomputed: pts[w * stride + i] = 2^(c w) P_i, w < W
struct G1Table {
    G1Affine* pts = nullptr;
    size_t n = 0, stride = 0;
    uint32_t c = 0, W = 0;
};
// window = 0: chosen from n (20 bits at 2^21: few digits, many buckets -- right for dense scalars);
// a smaller window trades digits for buckets (sparse / small scalars: the reduction over empty buckets dominates)
int g1_table_build(nzcb_ctx* ctx, const G1Affine* d_bases, size_t n, G1Table* out, uint32_t window = 0);
void g1_table_free(G1Table* t);
// K <= 4 MSMs over the first n[k] bases of one table as a single batch; results in d_out[0..K)
// sparse: most scalars are tiny (round 1 in the Lagrange basis): skip the batched-affine halving rounds
int msm_table_dev(nzcb_ctx* ctx, const G1Table& tab, const uint32_t* const* d_scalars, const size_t* n, int K,
                  bool scalars_mont, G1XYZZ* d_out, bool sparse = false);
// g1fft.cu : [L_i(tau)]G1 (i < 2^log_n) from [tau^j]G1 (j < 2^log_n); device buffers, in != out
int g1_lagrange_basis(nzcb_ctx* ctx, const G1Affine* d_srs, uint32_t log_n, G1Affine* d_out);
// D2H + stream sync + affine conversion of `count` <= 4 results
int msm_to_host_affine(nzcb_ctx* ctx, const G1XYZZ* d_pt, G1Affine* h_out, int count = 1);
// the same for results of msm_table_dev: in latency mode the ranks' partial sums are exchanged and added first
int msm_table_finish(nzcb_ctx* ctx, const G1XYZZ* d_pt, G1Affine* h_out, int count);
// destroys the latency mode's NCCL communicator of a context that is being freed (no-op otherwise)
void msm_split_release(nzcb_ctx* ctx);

}  // namespace nzcb
