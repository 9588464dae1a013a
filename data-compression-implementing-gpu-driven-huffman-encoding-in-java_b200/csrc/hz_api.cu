// hz_api.cu — context management and the stage-level C ABI of libhuffb200 (include/huffb200.h).
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "hz_common.cuh"

// ---- errors / utilities -----------------------------------------------------------------------
int hz_fail(hz_ctx* ctx, int code, const char* fmt, ...) {
    if (ctx) {
        char buf[512];
        va_list ap; va_start(ap, fmt);
        vsnprintf(buf, sizeof(buf), fmt, ap);
        va_end(ap);
        ctx->err = buf;
    }
    return code;
}

int hz_cuda_fail(hz_ctx* ctx, cudaError_t e, const char* what) {
    return hz_fail(ctx, e == cudaErrorMemoryAllocation ? HZ_ERR_NOMEM : HZ_ERR_CUDA,
                   "CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
}

int hz_reserve(hz_ctx* ctx, DevBuf* b, size_t bytes) {
    if (bytes == 0) bytes = 16;
    if (b->cap >= bytes) return HZ_OK;
    if (b->p) { cudaStreamSynchronize(ctx->stream); cudaFree(b->p); b->p = nullptr; b->cap = 0; }
    size_t want = bytes + bytes / 8 + 256;          // grow with slack; buffers are reused across calls
    cudaError_t e = cudaMalloc(&b->p, want);
    if (e != cudaSuccess) { cudaGetLastError(); e = cudaMalloc(&b->p, want = bytes); }
    if (e != cudaSuccess) return hz_cuda_fail(ctx, e, "cudaMalloc(scratch)");
    b->cap = want;
    return HZ_OK;
}

bool hz_is_device_ptr(const void* p) {
    if (!p) return false;
    cudaPointerAttributes a;
    cudaError_t e = cudaPointerGetAttributes(&a, p);
    if (e != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

static cudaEvent_t prof_event(hz_ctx* ctx) {
    if (!ctx->ev_pool.empty()) { cudaEvent_t e = ctx->ev_pool.back(); ctx->ev_pool.pop_back(); return e; }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}
void hz_prof_begin(hz_ctx* ctx) {
    if (!ctx->prof) return;
    ctx->ev_open = prof_event(ctx);
    cudaEventRecord(ctx->ev_open, ctx->stream);
}
void hz_prof_end(hz_ctx* ctx, const char* name) {
    if (!ctx->prof || !ctx->ev_open) return;
    cudaEvent_t e1 = prof_event(ctx);
    cudaEventRecord(e1, ctx->stream);
    ctx->prof_pending.push_back({name, ctx->ev_open, e1});
    ctx->ev_open = nullptr;
}
// Turn recorded event pairs into per-kernel totals (waits for the last recorded event).
void hz_prof_resolve(hz_ctx* ctx) {
    for (auto& p : ctx->prof_pending) {
        float ms = 0;
        cudaEventSynchronize(p.e1);
        if (cudaEventElapsedTime(&ms, p.e0, p.e1) != cudaSuccess) { cudaGetLastError(); ms = 0; }
        bool found = false;
        for (auto& e : ctx->prof_entries)
            if (strcmp(e.name, p.name) == 0) { e.ms += ms; e.launches++; found = true; break; }
        if (!found) ctx->prof_entries.push_back({p.name, (double)ms, 1});
        ctx->ev_pool.push_back(p.e0);
        ctx->ev_pool.push_back(p.e1);
    }
    ctx->prof_pending.clear();
}

void hz_read_knobs(hz_knobs* k) {
    auto num = [](const char* name, int dflt) { const char* ev = getenv(name); return ev ? atoi(ev) : dflt; };
    k->range_mult = num("HZ_RANGE_MULT", 0);
    if (const char* ev = getenv("HZ_HIST")) k->hist = strcmp(ev, "private") == 0 ? 0 : (strcmp(ev, "atomic") == 0 ? 1 : 2);
    k->hist_range = num("HZ_HIST_RANGE", 1);
    if (const char* ev = getenv("HZ_CODEBOOK")) k->codebook = strcmp(ev, "warp") == 0 ? 1 : 2;
    if (const char* ev = getenv("HZ_CODEBOOK_REPLAY")) k->codebook_lane0 = strcmp(ev, "lane0") == 0;
    k->ident = num("HZ_IDENT", 1) != 0;
    if (const char* ev = getenv("HZ_DEC")) k->dec_mode = strcmp(ev, "legacy") == 0 ? 1 : (strcmp(ev, "fused") == 0 ? 2 : 0);
    k->dec_prebuild = num("HZ_DEC_PREBUILD", -1);
    k->dec_win = num("HZ_DEC_WIN", 0);
    k->dec_groups = num("HZ_DEC_GROUPS", 0);
    k->dec_bulk = num("HZ_DEC_BULK", 1) != 0;
    k->fu_lead = num("HZ_FU_LEAD", 0);
    k->fu_grid = num("HZ_FU_GRID", 0);
    k->fu_warps = num("HZ_FU_WARPS", 0);
    k->enc_chain = num("HZ_ENC_CHAIN", 1);
    k->fu_cluster = num("HZ_FU_CLUSTER", -1);
    if (const char* ev = getenv("HZ_FU_DUMP")) k->fu_dump = ev;
}

static int check_status(hz_ctx* ctx) {
    // copy + reset the device status word; caller has synchronised or will right here
    cudaError_t e = cudaMemcpyAsync(ctx->h_status, ctx->d_status, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) return hz_cuda_fail(ctx, e, "status readback");
    int st = *ctx->h_status;
    if (st != 0) {
        cudaMemsetAsync(ctx->d_status, 0, sizeof(int), ctx->stream);
        return hz_fail(ctx, st, "device-side error: %s", hz_strerror(st));
    }
    return HZ_OK;
}

extern "C" {

const char* hz_strerror(int s) {
    switch (s) {
        case HZ_OK: return "ok";
        case HZ_ERR_ARG: return "bad argument";
        case HZ_ERR_CUDA: return "CUDA error";
        case HZ_ERR_NOMEM: return "out of memory";
        case HZ_ERR_CODE_TOO_LONG: return "a chunk needs a Huffman code longer than 32 bits";
        case HZ_ERR_OUT_TOO_SMALL: return "output buffer too small";
        case HZ_ERR_DECODE: return "Huffman decode error: bit pattern matches no codeword";
        case HZ_ERR_BAD_LENGTHS: return "invalid code-length table";
        case HZ_ERR_IO: return "I/O error";
        case HZ_ERR_FORMAT: return "Invalid file format";
        case HZ_ERR_CHECKSUM: return "Checksum mismatch";
        case HZ_ERR_UNSUPPORTED: return "not supported";
        default: return "unknown error";
    }
}

uint32_t hz_version(void) { return 0x000100; }

int hz_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

uint64_t hz_num_chunks(uint64_t n, uint32_t chunk_bytes) {
    return chunk_bytes ? (n + chunk_bytes - 1) / chunk_bytes : 0;
}

int hz_create(int device, hz_ctx** out_ctx) {
    if (!out_ctx) return HZ_ERR_ARG;
    *out_ctx = nullptr;
    int ndev = hz_device_count();
    if (ndev <= 0 || device < 0 || device >= ndev) return HZ_ERR_CUDA;   // no CPU fallback
    hz_ctx* c = new hz_ctx();
    c->device = device;
    hz_read_knobs(&c->knobs);
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaMalloc(&c->d_status, sizeof(int));
    if (e == cudaSuccess) e = cudaMemset(c->d_status, 0, sizeof(int));
    if (e == cudaSuccess) e = cudaHostAlloc(&c->h_status, sizeof(int), cudaHostAllocDefault);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device);
    if (e != cudaSuccess) { cudaGetLastError(); hz_destroy(c); return HZ_ERR_CUDA; }
    c->stream = c->own_stream;
    *out_ctx = c;
    return HZ_OK;
}

void hz_destroy(hz_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    hz_comm_destroy(c);
    DevBuf* bufs[] = {&c->glob, &c->seg_hist, &c->chunk_hist, &c->len, &c->code, &c->chunk_bits, &c->comp_size, &c->comp_off,
                      &c->seg_bitoff, &c->counter, &c->chain, &c->stage_in, &c->stage_out, &c->stage_a, &c->stage_b,
                      &c->stage_c, &c->stage_d, &c->stage_e, &c->dec_meta, &c->dec_rec, &c->dec_seqcnt, &c->dec_misc, &c->dec_tables};
    for (DevBuf* b : bufs) if (b->p) cudaFree(b->p);
    if (c->d_status) cudaFree(c->d_status);
    if (c->h_status) cudaFreeHost(c->h_status);
    if (c->h_pin) cudaFreeHost(c->h_pin);
    if (c->h_totals) cudaFreeHost(c->h_totals);
    for (int i = 0; i < hz_ctx::PIPE_SLOTS; ++i) {
        if (c->pipe_in[i].p) cudaFree(c->pipe_in[i].p);
        if (c->pipe_out[i].p) cudaFree(c->pipe_out[i].p);
        if (c->ev_in[i]) cudaEventDestroy(c->ev_in[i]);
        if (c->ev_comp[i]) cudaEventDestroy(c->ev_comp[i]);
        if (c->ev_out[i]) cudaEventDestroy(c->ev_out[i]);
    }
    DevBuf* pm[] = {&c->pipe_meta_a, &c->pipe_meta_b, &c->pipe_meta_c, &c->pipe_meta_d};
    for (DevBuf* b : pm) if (b->p) cudaFree(b->p);
    if (c->copy_in) cudaStreamDestroy(c->copy_in);
    if (c->copy_out) cudaStreamDestroy(c->copy_out);
    hz_prof_resolve(c);
    for (cudaEvent_t e : c->ev_pool) cudaEventDestroy(e);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    delete c;
}

const char* hz_last_error(const hz_ctx* ctx) { return ctx ? ctx->err.c_str() : "no context"; }

int hz_set_stream(hz_ctx* ctx, void* s) {
    if (!ctx) return HZ_ERR_ARG;
    ctx->stream = s ? (cudaStream_t)s : ctx->own_stream;
    return HZ_OK;
}

int hz_sync(hz_ctx* ctx) {
    if (!ctx) return HZ_ERR_ARG;
    HZ_CUDA(ctx, cudaSetDevice(ctx->device));
    return check_status(ctx);
}

int hz_prof_enable(hz_ctx* ctx, int on) { if (!ctx) return HZ_ERR_ARG; ctx->prof = on != 0; return HZ_OK; }
int hz_prof_reset(hz_ctx* ctx) {
    if (!ctx) return HZ_ERR_ARG;
    hz_prof_resolve(ctx);
    ctx->prof_entries.clear();
    return HZ_OK;
}
int hz_prof_count(hz_ctx* ctx) {
    if (!ctx) return 0;
    hz_prof_resolve(ctx);
    return (int)ctx->prof_entries.size();
}
int hz_prof_get(hz_ctx* ctx, int i, const char** name, double* total_ms, uint64_t* launches) {
    if (!ctx || i < 0 || i >= (int)ctx->prof_entries.size()) return HZ_ERR_ARG;
    if (name) *name = ctx->prof_entries[i].name;
    if (total_ms) *total_ms = ctx->prof_entries[i].ms;
    if (launches) *launches = ctx->prof_entries[i].launches;
    return HZ_OK;
}
uint64_t hz_launch_count(const hz_ctx* ctx) { return ctx ? ctx->launches : 0; }

}  // extern "C"

// ---- host <-> device staging helpers ---------------------------------------------------------
// Returns a device pointer holding `bytes` bytes of `p`: p itself when it is device memory, else
// a copy in scratch buffer `b`.
static int in_dev(hz_ctx* ctx, DevBuf* b, const void* p, size_t bytes, const void** d) {
    if (bytes == 0 || hz_is_device_ptr(p)) { *d = p; return HZ_OK; }
    HZ_TRY(hz_reserve(ctx, b, bytes));
    HZ_CUDA(ctx, cudaMemcpyAsync(b->p, p, bytes, cudaMemcpyHostToDevice, ctx->stream));
    *d = b->p;
    return HZ_OK;
}
// Device pointer to produce `bytes` bytes destined for `p` (p itself if device memory).
static int out_dev(hz_ctx* ctx, DevBuf* b, void* p, size_t bytes, void** d, bool* staged) {
    *staged = false;
    if (!p) { *d = nullptr; return HZ_OK; }
    if (hz_is_device_ptr(p)) { *d = p; return HZ_OK; }
    HZ_TRY(hz_reserve(ctx, b, bytes));
    *d = b->p; *staged = true;
    return HZ_OK;
}
static int out_copy(hz_ctx* ctx, void* host, const void* dev, size_t bytes) {
    if (bytes) HZ_CUDA(ctx, cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    return HZ_OK;
}

__global__ void sum_seg_hist_kernel(const uint32_t* __restrict__ seg_hist, uint32_t spc, uint32_t* __restrict__ hist) {
    const uint32_t k = blockIdx.x, t = threadIdx.x;
    const uint32_t* sh = seg_hist + (size_t)k * spc * 256 + t;
    uint32_t f = 0;
    for (uint32_t s = 0; s < spc; ++s) f += sh[(size_t)s * 256];
    hist[(size_t)k * 256 + t] = f;
}

extern "C" {

int hz_histogram(hz_ctx* ctx, const uint8_t* in, uint64_t n, uint32_t chunk_bytes, uint32_t* hist) {
    if (!ctx || !hist || chunk_bytes == 0 || (n && !in)) return hz_fail(ctx, HZ_ERR_ARG, "hz_histogram: bad argument");
    HZ_CUDA(ctx, cudaSetDevice(ctx->device));
    const uint64_t K64 = hz_num_chunks(n, chunk_bytes);
    if (K64 == 0) return HZ_OK;
    if (K64 > 0x7fffffffull) return hz_fail(ctx, HZ_ERR_ARG, "too many chunks");
    const uint32_t K = (uint32_t)K64;
    const uint32_t spc = (chunk_bytes + HZ_SEG_BYTES - 1) / HZ_SEG_BYTES;
    const void* d_in; void* d_hist; bool st;
    HZ_TRY(in_dev(ctx, &ctx->stage_in, in, n, &d_in));
    HZ_TRY(out_dev(ctx, &ctx->stage_a, hist, (size_t)K * 1024, &d_hist, &st));
    HZ_TRY(hz_reserve(ctx, &ctx->seg_hist, (size_t)K * spc * 1024));
    HZ_TRY(hzk_histogram(ctx, (const uint8_t*)d_in, n, chunk_bytes, K, (uint32_t*)ctx->seg_hist.p));
    HZ_LAUNCH(ctx, "sum_seg_hist", sum_seg_hist_kernel, K, 256, 0, (const uint32_t*)ctx->seg_hist.p, spc, (uint32_t*)d_hist);
    if (st) { HZ_TRY(out_copy(ctx, hist, d_hist, (size_t)K * 1024)); return check_status(ctx); }
    return HZ_OK;
}

int hz_build_codebooks(hz_ctx* ctx, const uint32_t* hist, uint32_t K, uint8_t* len, uint32_t* code) {
    if (!ctx || !hist || !len) return hz_fail(ctx, HZ_ERR_ARG, "hz_build_codebooks: bad argument");
    HZ_CUDA(ctx, cudaSetDevice(ctx->device));
    if (K == 0) return HZ_OK;
    const void* d_hist; void *d_len, *d_code; bool s1, s2;
    HZ_TRY(in_dev(ctx, &ctx->stage_a, hist, (size_t)K * 1024, &d_hist));
    HZ_TRY(out_dev(ctx, &ctx->stage_b, len, (size_t)K * 256, &d_len, &s1));
    HZ_TRY(out_dev(ctx, &ctx->stage_c, code, (size_t)K * 1024, &d_code, &s2));
    HZ_TRY(hzk_codebook(ctx, (const uint32_t*)d_hist, 0, K, nullptr, (uint8_t*)d_len, (uint32_t*)d_code,
                        nullptr, nullptr, nullptr, nullptr, nullptr));
    if (s1) HZ_TRY(out_copy(ctx, len, d_len, (size_t)K * 256));
    if (s2) HZ_TRY(out_copy(ctx, code, d_code, (size_t)K * 1024));
    if (s1 || s2) return check_status(ctx);
    return HZ_OK;
}

int hz_codes_from_lengths(hz_ctx* ctx, const uint8_t* len, uint32_t K, uint32_t* code) {
    if (!ctx || !len || !code) return hz_fail(ctx, HZ_ERR_ARG, "hz_codes_from_lengths: bad argument");
    HZ_CUDA(ctx, cudaSetDevice(ctx->device));
    if (K == 0) return HZ_OK;
    const void* d_len; void* d_code; bool st;
    HZ_TRY(in_dev(ctx, &ctx->stage_b, len, (size_t)K * 256, &d_len));
    HZ_TRY(out_dev(ctx, &ctx->stage_c, code, (size_t)K * 1024, &d_code, &st));
    HZ_TRY(hzk_codes_from_lengths(ctx, (const uint8_t*)d_len, K, (uint32_t*)d_code));
    if (st) { HZ_TRY(out_copy(ctx, code, d_code, (size_t)K * 1024)); return check_status(ctx); }
    return HZ_OK;
}

// ---- the pipeline of host buffers -------------------------------------------------------------
// hz_encode / hz_decode called with HOST input and output move the data in batches of whole chunks
// through a ring of PIPE_SLOTS device slots: the H2D copy of batch b+1, the kernels of batch b and
// the D2H copy of batch b-1 run concurrently on three streams (PCIe is full duplex).  Pinned host
// memory (hz_host_alloc) is needed for the copies to be asynchronous; pageable memory still works,
// serialised by the driver.
#define HZ_PIPE_BATCH_BYTES (64ull << 20)
#define HZ_PIPE_MIN_BYTES (128ull << 20)

static int pipe_init(hz_ctx* ctx) {
    if (ctx->copy_in) return HZ_OK;
    HZ_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_in, cudaStreamNonBlocking));
    HZ_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_out, cudaStreamNonBlocking));
    for (int i = 0; i < hz_ctx::PIPE_SLOTS; ++i) {
        HZ_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_in[i], cudaEventDisableTiming));
        HZ_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_comp[i], cudaEventDisableTiming));
        HZ_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_out[i], cudaEventDisableTiming));
    }
    return HZ_OK;
}
struct PipeQuiesce {                           // joins the copy streams when a pipelined call returns, also on errors
    hz_ctx* ctx;
    ~PipeQuiesce() {
        if (ctx->copy_in) cudaStreamSynchronize(ctx->copy_in);
        if (ctx->copy_out) cudaStreamSynchronize(ctx->copy_out);
        cudaStreamSynchronize(ctx->stream);
        cudaGetLastError();
    }
};
static int pipe_totals(hz_ctx* ctx, size_t n) {
    if (ctx->h_totals_cap >= n) return HZ_OK;
    if (ctx->h_totals) { cudaFreeHost(ctx->h_totals); ctx->h_totals = nullptr; ctx->h_totals_cap = 0; }
    cudaError_t e = cudaHostAlloc((void**)&ctx->h_totals, n * sizeof(uint64_t), cudaHostAllocMapped);
    if (e != cudaSuccess) return hz_cuda_fail(ctx, e, "cudaHostAlloc(totals)");
    ctx->h_totals_cap = n;
    return HZ_OK;
}
// the batch's payload size, written straight into pinned host memory (no copy-engine queueing)
__global__ void store_total_kernel(const uint64_t* __restrict__ d_total, uint64_t* __restrict__ h_slot) {
    *h_slot = *d_total;
    __threadfence_system();
}

// histogram -> codebook -> encode of K chunks, everything device-resident
// global != nullptr: ONE codebook from the histogram summed over all chunks (and, after hz_comm_init, over all
// ranks: hzk_global_histogram); its 256 lengths are also written to global_len256 (device)
static int encode_device(hz_ctx* ctx, const uint8_t* d_in, uint64_t n, uint32_t chunk_bytes, uint32_t K, const uint8_t* d_fixed,
                         uint8_t* d_out, uint64_t dcap, uint64_t* d_off, uint8_t* d_len, uint32_t* d_hist,
                         uint8_t* global_len256 = nullptr) {
    const uint32_t spc = (chunk_bytes + HZ_SEG_BYTES - 1) / HZ_SEG_BYTES;
    const size_t nseg = (size_t)K * spc;
    HZ_TRY(hz_reserve(ctx, &ctx->seg_hist, nseg * 1024));
    HZ_TRY(hz_reserve(ctx, &ctx->code, (size_t)K * 1024));
    HZ_TRY(hz_reserve(ctx, &ctx->chunk_bits, (size_t)K * 8));
    HZ_TRY(hz_reserve(ctx, &ctx->comp_size, (size_t)K * 4));
    HZ_TRY(hz_reserve(ctx, &ctx->seg_bitoff, nseg * 8));
    // Streams of >= 8 large chunks: histogram, codebooks and offsets in ONE launch whose CTAs build a chunk's codebook as
    // soon as its histogram is complete, and an encoder launched with programmatic stream serialization that waits per
    // chunk (hz_codebook.cu: hist_chain_kernel) - the codebook stage's 0.16 ms of latency leaves the critical path.
    // (Streams of smaller chunks and caller-supplied lengths keep the separate launches.)
    // Chunks of >= 2 MiB: a tail holds its CTA for ~0.16 ms, so 0.16 ms x 5.3 TB/s / chunk bytes tails are resident beside
    // the histogram at any time (16 MiB chunks: 50 of the 888 CTA slots, 4 MiB: 200).  Measured per 4 GiB: 16 MiB chunks
    // 3.17 -> 3.06 ms, 8 MiB 3.17 -> 3.10, 4 MiB 3.31 -> 3.20; per 1 GiB: 8 MiB 0.938 -> 0.862, 4 MiB 0.940 -> 0.884,
    // 2 MiB 0.962 -> 0.939 (per 4 GiB: 3 MiB 3.53 -> 3.37, 2 MiB 3.69 -> 3.53).  (With offsets chained prefix[k-1] -> prefix[k] instead of the decoupled look-back, 4 MiB
    // chunks were SLOWER, 3.30 -> 3.94 ms: ~2 us per hop in series against tails that finish 0.75 us apart.)
    const uint32_t chain_min = ctx->knobs.enc_chain > 1 ? (uint32_t)ctx->knobs.enc_chain << 10 : (2u << 20);   // developer knob: HZ_ENC_CHAIN=<min chunk KiB>
    if (ctx->knobs.enc_chain && !global_len256 && !d_fixed && K >= 8 && chunk_bytes >= chain_min && !(K >= 1024 && spc <= 32)) {
        const size_t kk = ((size_t)K + 1) & ~(size_t)1;
        const size_t bytes = 16 + kk * 4 + 2 * (size_t)K * 8;
        HZ_TRY(hz_reserve(ctx, &ctx->chain, bytes));
        uint8_t* m = (uint8_t*)ctx->chain.p;
        HzChain c;
        c.ticket = (uint32_t*)m; m += 16;
        c.prefix = (uint64_t*)m; m += (size_t)K * 8;
        c.ready = (uint64_t*)m; m += (size_t)K * 8;
        c.done = (uint32_t*)m;
        HZ_CUDA(ctx, cudaMemsetAsync(ctx->chain.p, 0, bytes, ctx->stream));
        HZ_TRY(hzk_hist_codebook_chain(ctx, d_in, n, chunk_bytes, K, (uint32_t*)ctx->seg_hist.p, c, d_hist, d_len,
                                       (uint32_t*)ctx->code.p, (uint64_t*)ctx->chunk_bits.p, (uint32_t*)ctx->comp_size.p,
                                       d_off, (uint64_t*)ctx->seg_bitoff.p));
        HZ_TRY(hzk_encode(ctx, d_in, n, chunk_bytes, K, d_len, (const uint32_t*)ctx->code.p, d_off,
                          (const uint64_t*)ctx->seg_bitoff.p, d_out, dcap, c.ready, (const uint32_t*)ctx->comp_size.p));
        return HZ_OK;
    }
    HZ_TRY(hzk_histogram(ctx, d_in, n, chunk_bytes, K, (uint32_t*)ctx->seg_hist.p));
    if (global_len256) {
        HZ_TRY(hz_reserve(ctx, &ctx->glob, 256 * 8 + 256 * 4));
        uint64_t* g64 = (uint64_t*)ctx->glob.p;
        uint32_t* h32 = (uint32_t*)(g64 + 256);
        HZ_TRY(hzk_global_histogram(ctx, (const uint32_t*)ctx->seg_hist.p, nseg, g64, h32));
        HZ_TRY(hzk_codebook(ctx, h32, 0, 1, nullptr, global_len256, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr));
        d_fixed = global_len256;
    }
    HZ_TRY(hzk_codebook(ctx, (const uint32_t*)ctx->seg_hist.p, spc, K, d_hist, d_len,
                        (uint32_t*)ctx->code.p, (uint64_t*)ctx->chunk_bits.p, (uint32_t*)ctx->comp_size.p,
                        d_off, (uint64_t*)ctx->seg_bitoff.p, d_fixed));
    HZ_TRY(hzk_encode(ctx, d_in, n, chunk_bytes, K, d_len, (const uint32_t*)ctx->code.p,
                      d_off, (const uint64_t*)ctx->seg_bitoff.p, d_out, dcap));
    return HZ_OK;
}

static int encode_pipelined(hz_ctx* ctx, const uint8_t* in, uint64_t n, uint32_t chunk_bytes, uint32_t K, const uint8_t* fixed_len,
                            uint8_t* out, uint64_t out_cap, uint64_t* comp_off, uint8_t* len_out, uint32_t* hist_out) {
    const int S = hz_ctx::PIPE_SLOTS;
    HZ_TRY(pipe_init(ctx));
    uint64_t per = HZ_PIPE_BATCH_BYTES / chunk_bytes;
    if (per == 0) per = 1;
    const uint32_t cpb = (uint32_t)per;                                // chunks per batch
    const uint32_t nb = (K + cpb - 1) / cpb;
    const size_t bbytes = (size_t)cpb * chunk_bytes;
    HZ_TRY(pipe_totals(ctx, nb));
    // a per-chunk Huffman code never expands a batch (mean length <= 8 bits); a caller-supplied table can, up to 32
    // bits per symbol: the output slots are then sized for that
    const size_t slot_cap = fixed_len ? 4 * bbytes : bbytes;
    for (int i = 0; i < S; ++i) {
        HZ_TRY(hz_reserve(ctx, &ctx->pipe_in[i], bbytes));
        HZ_TRY(hz_reserve(ctx, &ctx->pipe_out[i], slot_cap + 16));
    }
    HZ_TRY(hz_reserve(ctx, &ctx->pipe_meta_a, ((size_t)K + 1) * 8));   // batch-local offsets of every chunk
    HZ_TRY(hz_reserve(ctx, &ctx->pipe_meta_b, (size_t)K * 256));       // code lengths
    if (hist_out) HZ_TRY(hz_reserve(ctx, &ctx->pipe_meta_c, (size_t)K * 1024));
    const void* d_fixed = nullptr;
    if (fixed_len) {
        HZ_TRY(hz_reserve(ctx, &ctx->stage_e, 256));
        HZ_CUDA(ctx, cudaMemcpyAsync(ctx->stage_e.p, fixed_len, 256, cudaMemcpyHostToDevice, ctx->stream));
        d_fixed = ctx->stage_e.p;
    }
    uint64_t* d_off = (uint64_t*)ctx->pipe_meta_a.p;
    uint8_t* d_len = (uint8_t*)ctx->pipe_meta_b.p;
    uint32_t* d_hist = hist_out ? (uint32_t*)ctx->pipe_meta_c.p : nullptr;
    uint64_t* h_tot_dev = nullptr;
    HZ_CUDA(ctx, cudaHostGetDevicePointer((void**)&h_tot_dev, ctx->h_totals, 0));
    // make the pipeline streams start after whatever is already queued on the codec's stream
    HZ_CUDA(ctx, cudaEventRecord(ctx->ev_comp[0], ctx->stream));
    HZ_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_in, ctx->ev_comp[0], 0));
    HZ_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_out, ctx->ev_comp[0], 0));

    std::vector<uint64_t> base(nb + 1, 0);
    uint64_t drained = 0;                      // batches whose payload copy has been issued
    int rc = HZ_OK;
    PipeQuiesce quiesce{ctx};                  // every exit path waits for the copy streams: no DMA outlives the call
    auto drain = [&](uint32_t b) -> int {      // issue the D2H of batch b (its size is known once its kernels ran)
        const int s = b % S;
        HZ_CUDA(ctx, cudaEventSynchronize(ctx->ev_comp[s]));
        const uint64_t tot = ctx->h_totals[b];
        if (tot > slot_cap) return hz_fail(ctx, HZ_ERR_OUT_TOO_SMALL, "batch payload %llu > slot %llu", (unsigned long long)tot, (unsigned long long)slot_cap);
        base[b + 1] = base[b] + tot;
        if (base[b + 1] > out_cap) return hz_fail(ctx, HZ_ERR_OUT_TOO_SMALL, "payload > capacity %llu", (unsigned long long)out_cap);
        if (tot) HZ_CUDA(ctx, cudaMemcpyAsync(out + base[b], ctx->pipe_out[s].p, tot, cudaMemcpyDeviceToHost, ctx->copy_out));
        HZ_CUDA(ctx, cudaEventRecord(ctx->ev_out[s], ctx->copy_out));
        return HZ_OK;
    };
    for (uint32_t b = 0; b < nb && rc == HZ_OK; ++b) {
        const int s = b % S;
        const uint32_t k0 = b * cpb, kb = K - k0 < cpb ? K - k0 : cpb;
        const uint64_t o0 = (uint64_t)k0 * chunk_bytes, nbytes = n - o0 < bbytes ? n - o0 : bbytes;
        if (b >= (uint32_t)S) {
            // the slot is reused: its previous batch (b - S) must have been drained
            while (drained + S <= b && rc == HZ_OK) { rc = drain((uint32_t)drained); ++drained; }
            if (rc != HZ_OK) break;
            HZ_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_in, ctx->ev_comp[s], 0));      // kernels of b-S have read the input slot
            HZ_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_out[s], 0));       // payload of b-S has left the output slot
        }
        HZ_CUDA(ctx, cudaMemcpyAsync(ctx->pipe_in[s].p, in + o0, nbytes, cudaMemcpyHostToDevice, ctx->copy_in));
        HZ_CUDA(ctx, cudaEventRecord(ctx->ev_in[s], ctx->copy_in));
        HZ_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_in[s], 0));
        rc = encode_device(ctx, (const uint8_t*)ctx->pipe_in[s].p, nbytes, chunk_bytes, kb, (const uint8_t*)d_fixed,
                           (uint8_t*)ctx->pipe_out[s].p, slot_cap, d_off + k0, d_len + (size_t)k0 * 256,
                           d_hist ? d_hist + (size_t)k0 * 256 : nullptr);
        if (rc != HZ_OK) break;
        HZ_LAUNCH(ctx, "store_total", store_total_kernel, 1, 1, 0, d_off + k0 + kb, h_tot_dev + b);
        HZ_CUDA(ctx, cudaEventRecord(ctx->ev_comp[s], ctx->stream));
        // keep one batch of look-ahead: drain batch b-1 while batch b computes
        while (drained + 1 <= b && rc == HZ_OK) { rc = drain((uint32_t)drained); ++drained; }
    }
    while (rc == HZ_OK && drained < nb) { rc = drain((uint32_t)drained); ++drained; }
    if (rc == HZ_OK) {
        // metadata: batch-local offsets (rebased on the host), code lengths, histograms
        std::vector<uint64_t> loc((size_t)K + 1);
        HZ_CUDA(ctx, cudaMemcpyAsync(loc.data(), d_off, ((size_t)K + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream));
        if (len_out) HZ_CUDA(ctx, cudaMemcpyAsync(len_out, d_len, (size_t)K * 256, cudaMemcpyDeviceToHost, ctx->stream));
        if (hist_out) HZ_CUDA(ctx, cudaMemcpyAsync(hist_out, d_hist, (size_t)K * 1024, cudaMemcpyDeviceToHost, ctx->stream));
        rc = check_status(ctx);
        if (rc == HZ_OK) {
            for (uint32_t k = 0; k < K; ++k) comp_off[k] = base[k / cpb] + loc[k];
            comp_off[K] = base[nb];
        }
    }
    cudaError_t e = cudaStreamSynchronize(ctx->copy_out);
    cudaStreamSynchronize(ctx->copy_in);
    if (rc == HZ_OK && e != cudaSuccess) rc = hz_cuda_fail(ctx, e, "payload copy");
    return rc;
}

static int encode_impl(hz_ctx* ctx, const uint8_t* in, uint64_t n, uint32_t chunk_bytes, const uint8_t* fixed_len,
                       uint8_t* out, uint64_t out_cap, uint64_t* comp_off, uint8_t* len_out, uint32_t* hist_out) {
    if (!ctx || chunk_bytes == 0 || (n && (!in || !out)) || !comp_off)
        return hz_fail(ctx, HZ_ERR_ARG, "hz_encode: bad argument");
    HZ_CUDA(ctx, cudaSetDevice(ctx->device));
    const uint64_t K64 = hz_num_chunks(n, chunk_bytes);
    if (K64 > 0x7fffffffull) return hz_fail(ctx, HZ_ERR_ARG, "too many chunks");
    const uint32_t K = (uint32_t)K64;
    const bool off_dev = hz_is_device_ptr(comp_off);
    if (K == 0) {
        if (off_dev) HZ_CUDA(ctx, cudaMemsetAsync(comp_off, 0, sizeof(uint64_t), ctx->stream));
        else comp_off[0] = 0;
        return HZ_OK;
    }
    if (n >= HZ_PIPE_MIN_BYTES && !hz_is_device_ptr(in) && !hz_is_device_ptr(out) && !off_dev &&
        !(len_out && hz_is_device_ptr(len_out)) && !(hist_out && hz_is_device_ptr(hist_out)))
        return encode_pipelined(ctx, in, n, chunk_bytes, K, fixed_len, out, out_cap, comp_off, len_out, hist_out);
    const void *d_in, *d_fixed = nullptr;
    void *d_out, *d_off, *d_len, *d_hist;
    bool s_out, s_off, s_len, s_hist;
    HZ_TRY(in_dev(ctx, &ctx->stage_in, in, n, &d_in));
    if (fixed_len) HZ_TRY(in_dev(ctx, &ctx->stage_e, fixed_len, 256, &d_fixed));
    const uint64_t dcap = hz_is_device_ptr(out) ? out_cap : (fixed_len ? out_cap : (out_cap < n ? out_cap : n));
    HZ_TRY(out_dev(ctx, &ctx->stage_out, out, dcap + 16, &d_out, &s_out));
    HZ_TRY(out_dev(ctx, &ctx->comp_off, comp_off, ((size_t)K + 1) * 8, &d_off, &s_off));
    HZ_TRY(out_dev(ctx, &ctx->len, len_out, (size_t)K * 256, &d_len, &s_len));
    if (!d_len) { HZ_TRY(hz_reserve(ctx, &ctx->len, (size_t)K * 256)); d_len = ctx->len.p; }
    HZ_TRY(out_dev(ctx, &ctx->chunk_hist, hist_out, (size_t)K * 1024, &d_hist, &s_hist));
    HZ_TRY(encode_device(ctx, (const uint8_t*)d_in, n, chunk_bytes, K, (const uint8_t*)d_fixed, (uint8_t*)d_out, dcap,
                         (uint64_t*)d_off, (uint8_t*)d_len, (uint32_t*)d_hist));

    if (!(s_out || s_off || s_len || s_hist)) return HZ_OK;       // fully device-resident: asynchronous
    // host outputs: offsets first (their total tells how much payload to copy back)
    uint64_t total = 0;
    if (s_off) {
        HZ_TRY(out_copy(ctx, comp_off, d_off, ((size_t)K + 1) * 8));
    } else {
        HZ_CUDA(ctx, cudaMemcpyAsync(&total, (uint64_t*)d_off + K, 8, cudaMemcpyDeviceToHost, ctx->stream));
    }
    if (s_len) HZ_TRY(out_copy(ctx, len_out, d_len, (size_t)K * 256));
    if (s_hist) HZ_TRY(out_copy(ctx, hist_out, d_hist, (size_t)K * 1024));
    HZ_TRY(check_status(ctx));
    if (s_off) total = comp_off[K];
    if (s_out) {
        if (total > out_cap) return hz_fail(ctx, HZ_ERR_OUT_TOO_SMALL, "payload %llu > capacity %llu",
                                            (unsigned long long)total, (unsigned long long)out_cap);
        HZ_TRY(out_copy(ctx, out, d_out, total));
        HZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return HZ_OK;
}

int hz_encode(hz_ctx* ctx, const uint8_t* in, uint64_t n, uint32_t chunk_bytes, uint8_t* out, uint64_t out_cap,
              uint64_t* comp_off, uint8_t* len_out, uint32_t* hist_out) {
    return encode_impl(ctx, in, n, chunk_bytes, nullptr, out, out_cap, comp_off, len_out, hist_out);
}

int hz_encode_global(hz_ctx* ctx, const uint8_t* in, uint64_t n, uint32_t chunk_bytes, uint8_t* out, uint64_t out_cap,
                     uint64_t* comp_off, uint8_t* len256_out) {
    if (!ctx || chunk_bytes == 0 || (n && (!in || !out)) || !comp_off || !len256_out)
        return hz_fail(ctx, HZ_ERR_ARG, "hz_encode_global: bad argument");
    HZ_CUDA(ctx, cudaSetDevice(ctx->device));
    const uint64_t K64 = hz_num_chunks(n, chunk_bytes);
    if (K64 > 0x7fffffffull) return hz_fail(ctx, HZ_ERR_ARG, "too many chunks");
    const uint32_t K = (uint32_t)K64;
    // every rank takes part in the all-reduce, also one whose shard is empty (K == 0: an all-zero histogram)
    const void* d_in; void *d_out, *d_off, *d_gl; bool s_out, s_off, s_gl;
    HZ_TRY(in_dev(ctx, &ctx->stage_in, in, n, &d_in));
    HZ_TRY(out_dev(ctx, &ctx->stage_out, out, out_cap + 16, &d_out, &s_out));
    HZ_TRY(out_dev(ctx, &ctx->comp_off, comp_off, ((size_t)K + 1) * 8, &d_off, &s_off));
    HZ_TRY(out_dev(ctx, &ctx->stage_e, len256_out, 256, &d_gl, &s_gl));
    HZ_TRY(hz_reserve(ctx, &ctx->len, (size_t)(K ? K : 1) * 256));
    if (K == 0) {
        HZ_TRY(hz_reserve(ctx, &ctx->glob, 256 * 8 + 256 * 4));
        HZ_TRY(hz_reserve(ctx, &ctx->seg_hist, 1024));
        uint64_t* g64 = (uint64_t*)ctx->glob.p;
        HZ_TRY(hzk_global_histogram(ctx, (const uint32_t*)ctx->seg_hist.p, 0, g64, (uint32_t*)(g64 + 256)));
        HZ_TRY(hzk_codebook(ctx, (const uint32_t*)(g64 + 256), 0, 1, nullptr, (uint8_t*)d_gl, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr));
        HZ_CUDA(ctx, cudaMemsetAsync(d_off, 0, sizeof(uint64_t), ctx->stream));
    } else {
        HZ_TRY(encode_device(ctx, (const uint8_t*)d_in, n, chunk_bytes, K, nullptr, (uint8_t*)d_out, out_cap,
                             (uint64_t*)d_off, (uint8_t*)ctx->len.p, nullptr, (uint8_t*)d_gl));
    }
    if (!(s_out || s_off || s_gl)) return HZ_OK;
    uint64_t total = 0;
    if (s_off) HZ_TRY(out_copy(ctx, comp_off, d_off, ((size_t)K + 1) * 8));
    else HZ_CUDA(ctx, cudaMemcpyAsync(&total, (uint64_t*)d_off + K, 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (s_gl) HZ_TRY(out_copy(ctx, len256_out, d_gl, 256));
    HZ_TRY(check_status(ctx));
    if (s_off) total = comp_off[K];
    if (s_out) {
        if (total > out_cap) return hz_fail(ctx, HZ_ERR_OUT_TOO_SMALL, "payload %llu > capacity %llu",
                                            (unsigned long long)total, (unsigned long long)out_cap);
        HZ_TRY(out_copy(ctx, out, d_out, total));
        HZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return HZ_OK;
}

int hz_encode_with_lengths(hz_ctx* ctx, const uint8_t* in, uint64_t n, uint32_t chunk_bytes, const uint8_t* len256,
                           uint8_t* out, uint64_t out_cap, uint64_t* comp_off) {
    if (!len256) return hz_fail(ctx, HZ_ERR_ARG, "hz_encode_with_lengths: len256 is NULL");
    return encode_impl(ctx, in, n, chunk_bytes, len256, out, out_cap, comp_off, nullptr, nullptr);
}

static int decode_pipelined(hz_ctx* ctx, const uint8_t* comp, const uint64_t* comp_off, const uint32_t* comp_size,
                            const uint32_t* orig_size, const uint8_t* len, uint32_t K, uint8_t* out, uint64_t out_cap) {
    const int S = hz_ctx::PIPE_SLOTS;
    HZ_TRY(pipe_init(ctx));
    PipeQuiesce quiesce{ctx};                  // every exit path waits for the copy streams: no DMA outlives the call
    uint64_t per = HZ_PIPE_BATCH_BYTES / (orig_size[0] ? orig_size[0] : 1);
    if (per == 0) per = 1;
    const uint32_t cpb = (uint32_t)(per < K ? per : K);
    const uint32_t nb = (K + cpb - 1) / cpb;
    // batch geometry + batch-relative chunk offsets
    std::vector<uint64_t> rel(K), obase(nb + 1, 0), cbeg(nb), cend(nb);
    size_t max_in = 0, max_out = 0;
    for (uint32_t b = 0; b < nb; ++b) {
        const uint32_t k0 = b * cpb, k1 = k0 + cpb < K ? k0 + cpb : K;
        cbeg[b] = comp_off[k0]; cend[b] = comp_off[k1 - 1] + comp_size[k1 - 1];
        uint64_t o = 0;
        for (uint32_t k = k0; k < k1; ++k) { rel[k] = comp_off[k] - cbeg[b]; o += orig_size[k]; }
        obase[b + 1] = obase[b] + o;
        if (cend[b] - cbeg[b] > max_in) max_in = cend[b] - cbeg[b];
        if (o > max_out) max_out = o;
    }
    if (obase[nb] > out_cap) return hz_fail(ctx, HZ_ERR_OUT_TOO_SMALL, "decoded size %llu > capacity %llu",
                                            (unsigned long long)obase[nb], (unsigned long long)out_cap);
    for (int i = 0; i < S; ++i) {
        HZ_TRY(hz_reserve(ctx, &ctx->pipe_in[i], max_in + 32));
        HZ_TRY(hz_reserve(ctx, &ctx->pipe_out[i], max_out + 16));
    }
    HZ_TRY(hz_reserve(ctx, &ctx->pipe_meta_a, (size_t)K * 8));
    HZ_TRY(hz_reserve(ctx, &ctx->pipe_meta_b, (size_t)K * 256));
    HZ_TRY(hz_reserve(ctx, &ctx->pipe_meta_c, (size_t)K * 4));
    HZ_TRY(hz_reserve(ctx, &ctx->pipe_meta_d, (size_t)K * 4));
    uint64_t* d_rel = (uint64_t*)ctx->pipe_meta_a.p;
    uint8_t* d_len = (uint8_t*)ctx->pipe_meta_b.p;
    uint32_t* d_csz = (uint32_t*)ctx->pipe_meta_c.p;
    uint32_t* d_osz = (uint32_t*)ctx->pipe_meta_d.p;
    HZ_CUDA(ctx, cudaMemcpyAsync(d_rel, rel.data(), (size_t)K * 8, cudaMemcpyHostToDevice, ctx->stream));
    HZ_CUDA(ctx, cudaMemcpyAsync(d_len, len, (size_t)K * 256, cudaMemcpyHostToDevice, ctx->stream));
    HZ_CUDA(ctx, cudaMemcpyAsync(d_csz, comp_size, (size_t)K * 4, cudaMemcpyHostToDevice, ctx->stream));
    HZ_CUDA(ctx, cudaMemcpyAsync(d_osz, orig_size, (size_t)K * 4, cudaMemcpyHostToDevice, ctx->stream));
    HZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));          // `rel` is a local (pageable) vector
    HZ_CUDA(ctx, cudaEventRecord(ctx->ev_comp[0], ctx->stream));
    HZ_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_in, ctx->ev_comp[0], 0));
    HZ_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_out, ctx->ev_comp[0], 0));
    for (uint32_t b = 0; b < nb; ++b) {
        const int s = b % S;
        const uint32_t k0 = b * cpb, kb = K - k0 < cpb ? K - k0 : cpb;
        const uint64_t cbytes = cend[b] - cbeg[b], obytes = obase[b + 1] - obase[b];
        if (b >= (uint32_t)S) {
            HZ_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_in, ctx->ev_comp[s], 0));      // kernels of b-S have read the input slot
            HZ_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_out[s], 0));       // output of b-S has left its slot
        }
        if (cbytes) HZ_CUDA(ctx, cudaMemcpyAsync(ctx->pipe_in[s].p, comp + cbeg[b], cbytes, cudaMemcpyHostToDevice, ctx->copy_in));
        HZ_CUDA(ctx, cudaEventRecord(ctx->ev_in[s], ctx->copy_in));
        HZ_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_in[s], 0));
        HZ_TRY(hzk_decode(ctx, (const uint8_t*)ctx->pipe_in[s].p, cbytes, d_rel + k0, d_csz + k0, d_osz + k0, nullptr,
                          d_len + (size_t)k0 * 256, kb, (uint8_t*)ctx->pipe_out[s].p, obytes));
        HZ_CUDA(ctx, cudaEventRecord(ctx->ev_comp[s], ctx->stream));
        HZ_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_out, ctx->ev_comp[s], 0));
        if (obytes) HZ_CUDA(ctx, cudaMemcpyAsync(out + obase[b], ctx->pipe_out[s].p, obytes, cudaMemcpyDeviceToHost, ctx->copy_out));
        HZ_CUDA(ctx, cudaEventRecord(ctx->ev_out[s], ctx->copy_out));
    }
    int rc = check_status(ctx);
    cudaError_t e = cudaStreamSynchronize(ctx->copy_out);
    cudaStreamSynchronize(ctx->copy_in);
    if (rc == HZ_OK && e != cudaSuccess) rc = hz_cuda_fail(ctx, e, "output copy");
    return rc;
}

int hz_decode(hz_ctx* ctx, const uint8_t* comp, uint64_t comp_bytes, const uint64_t* comp_off,
              const uint32_t* comp_size, const uint32_t* orig_size, const uint64_t* orig_off,
              const uint8_t* len, uint32_t K, uint8_t* out, uint64_t out_cap) {
    if (!ctx || (K && (!comp_off || !comp_size || !orig_size || !len)) || (comp_bytes && !comp) || (out_cap && !out))
        return hz_fail(ctx, HZ_ERR_ARG, "hz_decode: bad argument");
    HZ_CUDA(ctx, cudaSetDevice(ctx->device));
    if (K == 0) return HZ_OK;
    // host buffers, chunks dense and in order, output back to back: pipeline the copies
    if (K >= 2 && !orig_off && out_cap >= HZ_PIPE_MIN_BYTES && !hz_is_device_ptr(comp) && !hz_is_device_ptr(out) &&
        !hz_is_device_ptr(comp_off) && !hz_is_device_ptr(comp_size) && !hz_is_device_ptr(orig_size) && !hz_is_device_ptr(len)) {
        bool dense = true;
        for (uint32_t k = 0; k + 1 < K && dense; ++k) dense = comp_off[k + 1] == comp_off[k] + comp_size[k];
        if (dense && comp_off[K - 1] + comp_size[K - 1] <= comp_bytes)
            return decode_pipelined(ctx, comp, comp_off, comp_size, orig_size, len, K, out, out_cap);
    }
    const void *d_comp, *d_coff, *d_csz, *d_osz, *d_ooff = nullptr, *d_len;
    void* d_out; bool s_out;
    HZ_TRY(in_dev(ctx, &ctx->stage_in, comp, comp_bytes, &d_comp));
    HZ_TRY(in_dev(ctx, &ctx->stage_a, comp_off, (size_t)K * 8, &d_coff));
    HZ_TRY(in_dev(ctx, &ctx->stage_b, comp_size, (size_t)K * 4, &d_csz));
    HZ_TRY(in_dev(ctx, &ctx->stage_c, orig_size, (size_t)K * 4, &d_osz));
    if (orig_off) HZ_TRY(in_dev(ctx, &ctx->stage_d, orig_off, (size_t)K * 8, &d_ooff));
    HZ_TRY(in_dev(ctx, &ctx->len, len, (size_t)K * 256, &d_len));
    HZ_TRY(out_dev(ctx, &ctx->stage_out, out, out_cap, &d_out, &s_out));
    HZ_TRY(hzk_decode(ctx, (const uint8_t*)d_comp, comp_bytes, (const uint64_t*)d_coff, (const uint32_t*)d_csz,
                      (const uint32_t*)d_osz, (const uint64_t*)d_ooff, (const uint8_t*)d_len, K, (uint8_t*)d_out, out_cap));
    if (s_out) {
        HZ_TRY(out_copy(ctx, out, d_out, out_cap));
        return check_status(ctx);
    }
    return HZ_OK;
}

/* pinned host memory for asynchronous (pipelined) transfers */
int hz_host_alloc(void** p, size_t bytes) {
    if (!p) return HZ_ERR_ARG;
    cudaError_t e = cudaHostAlloc(p, bytes ? bytes : 1, cudaHostAllocDefault);
    if (e != cudaSuccess) { cudaGetLastError(); *p = nullptr; return HZ_ERR_NOMEM; }
    return HZ_OK;
}
void hz_host_free(void* p) { if (p) cudaFreeHost(p); }

int hz_sha256_chunks(hz_ctx* ctx, const uint8_t* in, uint64_t n, uint32_t chunk_bytes, uint8_t* digests) {
    if (!ctx || chunk_bytes == 0 || !digests || (n && !in)) return hz_fail(ctx, HZ_ERR_ARG, "hz_sha256_chunks: bad argument");
    HZ_CUDA(ctx, cudaSetDevice(ctx->device));
    const uint64_t K64 = hz_num_chunks(n, chunk_bytes);
    if (K64 == 0) return HZ_OK;
    if (K64 > 0x7fffffffull) return hz_fail(ctx, HZ_ERR_ARG, "too many chunks");
    const uint32_t K = (uint32_t)K64;
    const void* d_in; void* d_dig; bool st;
    HZ_TRY(in_dev(ctx, &ctx->stage_in, in, n, &d_in));
    HZ_TRY(out_dev(ctx, &ctx->stage_a, digests, (size_t)K * 32, &d_dig, &st));
    HZ_TRY(hzk_sha256(ctx, (const uint8_t*)d_in, n, chunk_bytes, K, (uint8_t*)d_dig));
    if (st) { HZ_TRY(out_copy(ctx, digests, d_dig, (size_t)K * 32)); return check_status(ctx); }
    return HZ_OK;
}

}  // extern "C"
