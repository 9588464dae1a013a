// hz_common.cuh — shared constants, context layout and device helpers of libhuffb200.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
#include "../../include/huffb200.h"

// ---------------------------------------------------------------------------------------------
// Geometry shared by the histogram and encode kernels.
//   segment = the unit one CTA histograms and later encodes; a chunk is split into
//   ceil(chunk_bytes / HZ_SEG_BYTES) segments.  7 * 8192: a whole number of encoder tiles, and
//   with 256 threads every thread sees at most 224 bytes of a segment, so the per-thread 8-bit
//   counters of the alternative histogram kernel cannot wrap.
// ---------------------------------------------------------------------------------------------
#define HZ_THREADS 256
#define HZ_SEG_BYTES 57344u
// An encoder group codes a RANGE of `mult` consecutive segments of a chunk from one starting bit offset
// (longer ranges amortise the per-range start-up; short chunks keep them short so that both groups of a
// CTA have work).  The histogram kernel counts a whole range per CTA into the range's FIRST segment slot
// and zeroes the others: every consumer either sums a chunk's slots or reads the offset of a range start.
static inline uint32_t hz_range_mult(uint32_t spc, int knob) {
    uint32_t m = spc >= 64 ? 4u : (spc >= 16 ? 2u : 1u);
    if (knob >= 1 && knob <= 64 && spc >= 64) m = (uint32_t)knob;
    return m;
}

// Developer knobs (A/B switches of kernel variants, never needed in production): read from the environment ONCE,
// when the context is created (hz_create), not on every launch.
struct hz_knobs {
    int range_mult = 0;        // HZ_RANGE_MULT   segments per encoder range (0 = default rule)
    int hist = 2;              // HZ_HIST         private | atomic | lanes (default)
    int hist_range = 1;        // HZ_HIST_RANGE   0 = one segment per histogram CTA
    int codebook = 0;          // HZ_CODEBOOK     0 = default rule, 1 = warp, 2 = lane
    int codebook_lane0 = 0;    // HZ_CODEBOOK_REPLAY=lane0
    int ident = 1;             // HZ_IDENT        0 = identity chunks go through the coder
    int dec_mode = 0;          // HZ_DEC          0 = by chunk size, 1 = legacy multi-pass kernels, 2 = fused kernel
    int dec_prebuild = -1;     // HZ_DEC_PREBUILD (legacy kernels)
    int dec_win = 0;           // HZ_DEC_WIN
    int dec_groups = 0;        // HZ_DEC_GROUPS
    int dec_bulk = 1;          // HZ_DEC_BULK
    int fu_lead = 0;           // HZ_FU_LEAD      lead-in words of the fused decoder
    int fu_grid = 0;           // HZ_FU_GRID      CTAs of the fused decoder
    int fu_warps = 0;          // HZ_FU_WARPS     warps per CTA of the fused decoder (24, 8 or 5)
    int enc_chain = 1;         // HZ_ENC_CHAIN    0: separate histogram / codebook / offsets / encode launches
    int fu_cluster = -1;       // HZ_FU_CLUSTER   0 / 1: never / always launch the 24-warp fused decoder as cluster pairs
    std::string fu_dump;       // HZ_FU_DUMP      file for per-subsequence records
};
void hz_read_knobs(hz_knobs* k);

struct hz_prof_entry { const char* name; double ms; uint64_t launches; };

// Stage timings of the last file-level call, in the order of the reference's StageMetrics.Stage (hz_stage_metrics)
struct StageAcc {
    double ms[HZ_STAGE_COUNT] = {}; uint64_t count[HZ_STAGE_COUNT] = {}, bytes[HZ_STAGE_COUNT] = {};
    void add(int stage, double milli, uint64_t nbytes) { ms[stage] += milli; count[stage]++; bytes[stage] += nbytes; }
};

struct DevBuf {                               // grow-only device scratch buffer
    void* p = nullptr; size_t cap = 0;
};

struct hz_ctx {
    int device = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    std::string err;
    uint64_t launches = 0;
    int sm_count = 148;
    hz_knobs knobs;
    StageAcc stages;
    void* nccl_comm = nullptr; int nccl_ranks = 1, nccl_rank = 0;     // global-codebook mode (hz_comm_init)
    DevBuf glob;                                                      // u64[256] + u32[256] + u8[256]
    // kernel attributes (opt-in shared memory) are per device: set once per context, not once per process
    bool attr_encode = false, attr_decode = false, attr_decode_fused = false, attr_hist = false, attr_codebook = false;
    // device-side status word (first error latched by kernels) + pinned host mirror
    int* d_status = nullptr;
    int* h_status = nullptr;
    // scratch
    DevBuf seg_hist, chunk_hist, len, code, chunk_bits, comp_size, comp_off, seg_bitoff, counter, chain;
    DevBuf stage_in, stage_out, stage_a, stage_b, stage_c, stage_d, stage_e;
    DevBuf dec_meta, dec_rec, dec_seqcnt, dec_misc, dec_tables;
    void* h_pin = nullptr; size_t h_pin_cap = 0;
    // host-buffer pipeline (hz_encode / hz_decode with host pointers): copy streams, ring of device slots
    static const int PIPE_SLOTS = 3;
    cudaStream_t copy_in = nullptr, copy_out = nullptr;
    cudaEvent_t ev_in[PIPE_SLOTS] = {}, ev_comp[PIPE_SLOTS] = {}, ev_out[PIPE_SLOTS] = {};
    DevBuf pipe_in[PIPE_SLOTS], pipe_out[PIPE_SLOTS], pipe_meta_a, pipe_meta_b, pipe_meta_c, pipe_meta_d;
    uint64_t* h_totals = nullptr; size_t h_totals_cap = 0;     // pinned, written by the device (zero-copy)
    // profiling
    // profiling: event pairs are recorded without synchronising and resolved lazily
    bool prof = false;
    std::vector<hz_prof_entry> prof_entries;
    struct PendingProf { const char* name; cudaEvent_t e0, e1; };
    std::vector<PendingProf> prof_pending;
    std::vector<cudaEvent_t> ev_pool;
    cudaEvent_t ev_open = nullptr;
};

int hz_fail(hz_ctx* ctx, int code, const char* fmt, ...);
int hz_cuda_fail(hz_ctx* ctx, cudaError_t e, const char* what);
int hz_reserve(hz_ctx* ctx, DevBuf* b, size_t bytes);
bool hz_is_device_ptr(const void* p);
void hz_prof_begin(hz_ctx* ctx);
void hz_prof_end(hz_ctx* ctx, const char* name);
void hz_prof_resolve(hz_ctx* ctx);

#define HZ_CUDA(ctx, call)                                                   \
    do {                                                                     \
        cudaError_t e__ = (call);                                            \
        if (e__ != cudaSuccess) return hz_cuda_fail((ctx), e__, #call);      \
    } while (0)

#define HZ_TRY(expr)                     \
    do {                                 \
        int rc__ = (expr);               \
        if (rc__ != HZ_OK) return rc__;  \
    } while (0)

// kernel launch wrapper: counts launches, optional per-kernel event timing
#define HZ_LAUNCH(ctx, name, kernel, grid, block, smem, ...)                              \
    do {                                                                                  \
        hz_prof_begin(ctx);                                                               \
        kernel<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);                  \
        (ctx)->launches++;                                                                \
        hz_prof_end(ctx, name);                                                           \
        cudaError_t e__ = cudaGetLastError();                                             \
        if (e__ != cudaSuccess) return hz_cuda_fail((ctx), e__, name);                    \
    } while (0)

// flags of the chained histogram -> codebook -> encode pipeline (zeroed before every chained call)
struct HzChain {
    uint32_t* ticket;       // [1]  next histogram range
    uint32_t* done;         // [K]  histogram ranges of chunk k that are complete
    uint64_t* prefix;       // [K]  top two bits: 0 empty, 1 aggregate (this chunk's size), 2 prefix (payload bytes of chunks 0..k)
    uint64_t* ready;        // [K]  != 0: chunk k's lengths, codes, offsets are in memory
};

// ---- launchers implemented in the kernel files ------------------------------------------------
int hzk_histogram(hz_ctx* ctx, const uint8_t* d_in, uint64_t n, uint32_t chunk_bytes, uint32_t K,
                  uint32_t* d_seg_hist);
int hzk_codebook(hz_ctx* ctx, const uint32_t* d_seg_hist, uint32_t segs_per_chunk, uint32_t K,
                 uint32_t* d_chunk_hist, uint8_t* d_len, uint32_t* d_code, uint64_t* d_chunk_bits,
                 uint32_t* d_comp_size, uint64_t* d_comp_off, uint64_t* d_seg_bitoff,
                 const uint8_t* d_fixed_len256);
int hzk_hist_codebook_chain(hz_ctx* ctx, const uint8_t* d_in, uint64_t n, uint32_t chunk_bytes, uint32_t K, uint32_t* d_seg_hist,
                            const HzChain& c, uint32_t* d_chunk_hist, uint8_t* d_len, uint32_t* d_code, uint64_t* d_chunk_bits,
                            uint32_t* d_comp_size, uint64_t* d_comp_off, uint64_t* d_seg_bitoff);
int hzk_codes_from_lengths(hz_ctx* ctx, const uint8_t* d_len, uint32_t K, uint32_t* d_code);
int hzk_encode(hz_ctx* ctx, const uint8_t* d_in, uint64_t n, uint32_t chunk_bytes, uint32_t K,
               const uint8_t* d_len, const uint32_t* d_code, const uint64_t* d_comp_off,
               const uint64_t* d_seg_bitoff, uint8_t* d_out, uint64_t out_cap,
               const uint64_t* d_chain_ready = nullptr, const uint32_t* d_comp_size = nullptr);
int hzk_decode(hz_ctx* ctx, const uint8_t* d_comp, uint64_t comp_bytes, const uint64_t* d_comp_off,
               const uint32_t* d_comp_size, const uint32_t* d_orig_size, const uint64_t* d_orig_off,
               const uint8_t* d_len, uint32_t K, uint8_t* d_out, uint64_t out_cap);
int hzk_decode_fused(hz_ctx* ctx, const uint8_t* d_comp, uint64_t comp_bytes, const uint64_t* d_comp_off,
                     const uint32_t* d_comp_size, const uint32_t* d_orig_size, const uint64_t* d_orig_off,
                     const uint8_t* d_len, uint32_t K, uint8_t* d_out, uint64_t out_cap, const uint8_t* d_ident,
                     const uint64_t** plan_orig_off, const uint32_t** plan_islice);
int hzk_global_histogram(hz_ctx* ctx, const uint32_t* d_seg_hist, uint64_t nseg, uint64_t* d_g64, uint32_t* d_h32);
int hzk_sha256(hz_ctx* ctx, const uint8_t* d_in, uint64_t n, uint32_t chunk_bytes, uint32_t K,
               uint8_t* d_digests);

#ifdef __CUDACC__
// ---- device helpers ---------------------------------------------------------------------------
__device__ __forceinline__ uint4 ld_stream_u4(const uint4* p) {   // streaming 128-bit load
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ uint32_t bswap32(uint32_t x) { return __byte_perm(x, 0, 0x0123); }

__device__ __forceinline__ void hz_set_status(int* status, int code) {
    if (code != 0) atomicCAS(status, 0, code);
}

// Copies n bytes src -> dst with the nthr threads of a group (tid = 0 .. nthr-1), any alignment of either
// pointer: dst-aligned 128-bit stores; a source that is not 16-byte aligned is read as 4-byte aligned words
// and funnel-shifted into place (an aligned word that holds one valid byte never leaves the allocation).
// Used where a chunk's code is the identity (all 256 symbols with 8-bit codes: canonical code == symbol),
// so that encoding / decoding it is a byte copy.
__device__ __forceinline__ void hz_group_copy(uint8_t* dst, const uint8_t* src, uint64_t n, uint32_t tid, uint32_t nthr) {
    uint64_t head = (16 - (reinterpret_cast<uintptr_t>(dst) & 15)) & 15;
    if (head > n) head = n;
    if (tid < head) dst[tid] = src[tid];
    const uint8_t* s = src + head;
    uint4* d4 = reinterpret_cast<uint4*>(dst + head);
    const uint64_t units = (n - head) >> 4;
    if ((reinterpret_cast<uintptr_t>(s) & 15) == 0) {
        const uint4* s4 = reinterpret_cast<const uint4*>(s);
#pragma unroll 4
        for (uint64_t u = tid; u < units; u += nthr) d4[u] = ld_stream_u4(s4 + u);
    } else {
        const uint32_t r = (uint32_t)(reinterpret_cast<uintptr_t>(s) & 3), sh = r * 8;
        const uint32_t* sw = reinterpret_cast<const uint32_t*>(s - r);
#pragma unroll 4
        for (uint64_t u = tid; u < units; u += nthr) {
            const uint32_t* w = sw + u * 4;
            const uint32_t w0 = __ldg(w), w1 = __ldg(w + 1), w2 = __ldg(w + 2), w3 = __ldg(w + 3);
            const uint32_t w4 = r ? __ldg(w + 4) : 0u;
            d4[u] = make_uint4(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh),
                               __funnelshift_r(w2, w3, sh), __funnelshift_r(w3, w4, sh));
        }
    }
    const uint64_t done = head + (units << 4);
    if (tid < n - done) dst[done + tid] = src[done + tid];
}
#endif
