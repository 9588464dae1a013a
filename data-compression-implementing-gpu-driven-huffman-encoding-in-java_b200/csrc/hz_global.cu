// hz_global.cu — the path's ONE collective: sum of the per-GPU byte histograms for the global-codebook
// extension mode (BASELINE.json config 4 "NCCL histogram allreduce").
//
// The reference has no multi-GPU code, so nothing is replaced here; the mode exists because one logical file
// sharded over G GPUs may want ONE codebook (a valid .dcz that any reference decoder accepts: the footer
// simply repeats the same 256 code lengths in every chunk record, core/CompressionHeader.java:80-83).
// NCCL is bound at run time (dlopen of libnccl.so.2 — the copy the host process already loaded, if any), so
// the library has no link-time dependency on it and single-GPU users never touch it.  The all-reduce is
// enqueued on the codec's stream: 256 x u64 = 2 KiB, latency bound, no host synchronisation.
#include <dlfcn.h>
#include <nccl.h>
#include "hz_common.cuh"

struct hz_nccl_api {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

static hz_nccl_api* nccl_api() {
    static hz_nccl_api api;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (h) {
            api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(h, "ncclGetUniqueId");
            api.CommInitRank = (decltype(api.CommInitRank))dlsym(h, "ncclCommInitRank");
            api.AllReduce = (decltype(api.AllReduce))dlsym(h, "ncclAllReduce");
            api.CommDestroy = (decltype(api.CommDestroy))dlsym(h, "ncclCommDestroy");
            api.GetErrorString = (decltype(api.GetErrorString))dlsym(h, "ncclGetErrorString");
            if (api.GetUniqueId && api.CommInitRank && api.AllReduce && api.CommDestroy) api.lib = h;
        }
    }
    return api.lib ? &api : nullptr;
}

static int nccl_fail(hz_ctx* ctx, hz_nccl_api* a, ncclResult_t r, const char* what) {
    return hz_fail(ctx, HZ_ERR_CUDA, "NCCL error %d (%s) in %s", (int)r, a && a->GetErrorString ? a->GetErrorString(r) : "?", what);
}

// sum of all segment histograms of this GPU's shard -> u64[256] (zeroed by the caller)
__global__ void __launch_bounds__(256)
global_hist_kernel(const uint32_t* __restrict__ seg_hist, uint64_t nseg, unsigned long long* __restrict__ g) {
    unsigned long long f = 0;
    for (uint64_t s = blockIdx.x; s < nseg; s += gridDim.x) f += seg_hist[s * 256 + threadIdx.x];
    if (f) atomicAdd(g + threadIdx.x, f);
}

// hz_build_codebooks takes 32-bit counts: scale the summed histogram down by a power of two when a bin needs more
// (16 GiB of a skewed stream do); a symbol that occurs keeps a count >= 1.  Every rank applies the same rule to
// the same all-reduced numbers, so every rank builds the same codebook.
__global__ void __launch_bounds__(256)
global_scale_kernel(const unsigned long long* __restrict__ g, uint32_t* __restrict__ h32) {
    __shared__ unsigned long long mx[256];
    const uint32_t t = threadIdx.x;
    const unsigned long long v = g[t];
    mx[t] = v;
    __syncthreads();
    for (int d = 128; d > 0; d >>= 1) { if (t < d && mx[t + d] > mx[t]) mx[t] = mx[t + d]; __syncthreads(); }
    const unsigned long long m = mx[0];
    const int s = m >> 31 ? 64 - __clzll((long long)m) - 31 : 0;
    unsigned long long r = s ? (v + (1ull << (s - 1))) >> s : v;
    if (v && !r) r = 1;
    h32[t] = (uint32_t)r;
}

int hzk_global_histogram(hz_ctx* ctx, const uint32_t* d_seg_hist, uint64_t nseg, uint64_t* d_g64, uint32_t* d_h32) {
    HZ_CUDA(ctx, cudaMemsetAsync(d_g64, 0, 256 * sizeof(uint64_t), ctx->stream));
    const unsigned grid = (unsigned)(nseg < 4096 ? (nseg ? nseg : 1) : 4096);
    HZ_LAUNCH(ctx, "global_hist", global_hist_kernel, grid, 256, 0, d_seg_hist, nseg, (unsigned long long*)d_g64);
    if (ctx->nccl_comm && ctx->nccl_ranks > 1) {
        hz_nccl_api* a = nccl_api();
        if (!a) return hz_fail(ctx, HZ_ERR_UNSUPPORTED, "libnccl.so.2 not found");
        hz_prof_begin(ctx);
        ncclResult_t r = a->AllReduce(d_g64, d_g64, 256, ncclUint64, ncclSum, (ncclComm_t)ctx->nccl_comm, ctx->stream);
        hz_prof_end(ctx, "nccl_allreduce_hist");
        if (r != ncclSuccess) return nccl_fail(ctx, a, r, "ncclAllReduce");
    }
    HZ_LAUNCH(ctx, "global_scale", global_scale_kernel, 1, 256, 0, (const unsigned long long*)d_g64, d_h32);
    return HZ_OK;
}

extern "C" {

int hz_comm_unique_id(void* id128) {
    hz_nccl_api* a = nccl_api();
    if (!a || !id128) return a ? HZ_ERR_ARG : HZ_ERR_UNSUPPORTED;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    return a->GetUniqueId((ncclUniqueId*)id128) == ncclSuccess ? HZ_OK : HZ_ERR_CUDA;
}

int hz_comm_init(hz_ctx* ctx, const void* id128, int nranks, int rank) {
    if (!ctx || !id128 || nranks < 1 || rank < 0 || rank >= nranks) return hz_fail(ctx, HZ_ERR_ARG, "hz_comm_init: bad argument");
    hz_nccl_api* a = nccl_api();
    if (!a) return hz_fail(ctx, HZ_ERR_UNSUPPORTED, "libnccl.so.2 not found");
    HZ_CUDA(ctx, cudaSetDevice(ctx->device));
    if (ctx->nccl_comm) { a->CommDestroy((ncclComm_t)ctx->nccl_comm); ctx->nccl_comm = nullptr; }
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    ncclComm_t comm = nullptr;
    ncclResult_t r = a->CommInitRank(&comm, nranks, id, rank);
    if (r != ncclSuccess) return nccl_fail(ctx, a, r, "ncclCommInitRank");
    ctx->nccl_comm = comm; ctx->nccl_ranks = nranks; ctx->nccl_rank = rank;
    return HZ_OK;
}

int hz_comm_destroy(hz_ctx* ctx) {
    if (!ctx) return HZ_ERR_ARG;
    if (ctx->nccl_comm) {
        if (hz_nccl_api* a = nccl_api()) a->CommDestroy((ncclComm_t)ctx->nccl_comm);
        ctx->nccl_comm = nullptr; ctx->nccl_ranks = 1; ctx->nccl_rank = 0;
    }
    return HZ_OK;
}

}  // extern "C"
