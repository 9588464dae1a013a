// hz_container.cpp — host side of the CompressionService surface: the .dcz container
// (core/CompressionHeader.java:51-144, core/ChunkMetadata.java:12-18, docs/FILE_FORMAT.md),
// SHA-256 (util/ChecksumUtil.java:11-27) and the file / buffer level compress, decompress and
// verify entry points that mirror CpuCompressionService.compress / decompress / verifyIntegrity
// (service/cpu/CpuCompressionService.java:57-205, :318-506, :644-696).
//
// The hot path (histogram -> codebook -> encode, and decode) runs on the GPU through the
// stage-level C ABI; this file only moves bytes, hashes them and lays out the footer.  Unlike the
// reference (which keeps every compressed chunk in RAM until the end, :81,160-163) the payload is
// streamed to the output batch by batch; the bytes written are identical.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cerrno>
#include <string>
#include <vector>
#include <thread>
#include <future>
#include <atomic>
#include <chrono>
#include <algorithm>
#include <fcntl.h>
#include <unistd.h>
#include <sys/stat.h>
#include <immintrin.h>
#include "hz_common.cuh"

// ================================================================================================
// SHA-256 (host).  SHA-NI when the CPU has it, portable FIPS 180-4 otherwise.
// ================================================================================================
namespace {

const uint32_t K256[64] = {
    0x428a2f98,0x71374491,0xb5c0fbcf,0xe9b5dba5,0x3956c25b,0x59f111f1,0x923f82a4,0xab1c5ed5,
    0xd807aa98,0x12835b01,0x243185be,0x550c7dc3,0x72be5d74,0x80deb1fe,0x9bdc06a7,0xc19bf174,
    0xe49b69c1,0xefbe4786,0x0fc19dc6,0x240ca1cc,0x2de92c6f,0x4a7484aa,0x5cb0a9dc,0x76f988da,
    0x983e5152,0xa831c66d,0xb00327c8,0xbf597fc7,0xc6e00bf3,0xd5a79147,0x06ca6351,0x14292967,
    0x27b70a85,0x2e1b2138,0x4d2c6dfc,0x53380d13,0x650a7354,0x766a0abb,0x81c2c92e,0x92722c85,
    0xa2bfe8a1,0xa81a664b,0xc24b8b70,0xc76c51a3,0xd192e819,0xd6990624,0xf40e3585,0x106aa070,
    0x19a4c116,0x1e376c08,0x2748774c,0x34b0bcb5,0x391c0cb3,0x4ed8aa4a,0x5b9cca4f,0x682e6ff3,
    0x748f82ee,0x78a5636f,0x84c87814,0x8cc70208,0x90befffa,0xa4506ceb,0xbef9a3f7,0xc67178f2};

inline uint32_t ror(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }

void sha_blocks_portable(uint32_t h[8], const uint8_t* p, size_t nblk) {
    for (; nblk; --nblk, p += 64) {
        uint32_t w[64];
        for (int i = 0; i < 16; ++i)
            w[i] = (uint32_t)p[4*i] << 24 | (uint32_t)p[4*i+1] << 16 | (uint32_t)p[4*i+2] << 8 | p[4*i+3];
        for (int i = 16; i < 64; ++i) {
            uint32_t s0 = ror(w[i-15], 7) ^ ror(w[i-15], 18) ^ (w[i-15] >> 3);
            uint32_t s1 = ror(w[i-2], 17) ^ ror(w[i-2], 19) ^ (w[i-2] >> 10);
            w[i] = w[i-16] + s0 + w[i-7] + s1;
        }
        uint32_t a=h[0],b=h[1],c=h[2],d=h[3],e=h[4],f=h[5],g=h[6],hh=h[7];
        for (int i = 0; i < 64; ++i) {
            uint32_t t1 = hh + (ror(e,6) ^ ror(e,11) ^ ror(e,25)) + ((e & f) ^ (~e & g)) + K256[i] + w[i];
            uint32_t t2 = (ror(a,2) ^ ror(a,13) ^ ror(a,22)) + ((a & b) ^ (a & c) ^ (b & c));
            hh=g; g=f; f=e; e=d+t1; d=c; c=b; b=a; a=t1+t2;
        }
        h[0]+=a;h[1]+=b;h[2]+=c;h[3]+=d;h[4]+=e;h[5]+=f;h[6]+=g;h[7]+=hh;
    }
}

__attribute__((target("sha,sse4.1,ssse3")))
void sha_blocks_ni(uint32_t h[8], const uint8_t* p, size_t nblk) {
    const __m128i MASK = _mm_set_epi64x(0x0c0d0e0f08090a0bULL, 0x0405060700010203ULL);
    __m128i tmp = _mm_loadu_si128((const __m128i*)&h[0]);
    __m128i st1 = _mm_loadu_si128((const __m128i*)&h[4]);
    tmp = _mm_shuffle_epi32(tmp, 0xB1);
    st1 = _mm_shuffle_epi32(st1, 0x1B);
    __m128i st0 = _mm_alignr_epi8(tmp, st1, 8);
    st1 = _mm_blend_epi16(st1, tmp, 0xF0);
    for (; nblk; --nblk, p += 64) {
        const __m128i a0 = st0, a1 = st1;
        __m128i m[4];
        for (int i = 0; i < 4; ++i) m[i] = _mm_shuffle_epi8(_mm_loadu_si128((const __m128i*)(p + 16 * i)), MASK);
#pragma GCC unroll 16
        for (int r = 0; r < 16; ++r) {
            // m[r&3] holds W[4r..4r+3]
            __m128i msg = _mm_add_epi32(m[r & 3], _mm_loadu_si128((const __m128i*)&K256[4 * r]));
            st1 = _mm_sha256rnds2_epu32(st1, st0, msg);
            if (r >= 3 && r <= 14) {     // finish W[4(r+1)..] (msg1 was applied two groups ago)
                __m128i tmp2 = _mm_alignr_epi8(m[r & 3], m[(r + 3) & 3], 4);
                m[(r + 1) & 3] = _mm_sha256msg2_epu32(_mm_add_epi32(m[(r + 1) & 3], tmp2), m[r & 3]);
            }
            st0 = _mm_sha256rnds2_epu32(st0, st1, _mm_shuffle_epi32(msg, 0x0E));
            if (r >= 1 && r <= 12)       // start W[4(r+3)..]
                m[(r + 3) & 3] = _mm_sha256msg1_epu32(m[(r + 3) & 3], m[r & 3]);
        }
        st0 = _mm_add_epi32(st0, a0);
        st1 = _mm_add_epi32(st1, a1);
    }
    tmp = _mm_shuffle_epi32(st0, 0x1B);
    st1 = _mm_shuffle_epi32(st1, 0xB1);
    st0 = _mm_blend_epi16(tmp, st1, 0xF0);
    st1 = _mm_alignr_epi8(st1, tmp, 8);
    _mm_storeu_si128((__m128i*)&h[0], st0);
    _mm_storeu_si128((__m128i*)&h[4], st1);
}

bool have_sha_ni() {
    static int v = -1;
    if (v < 0) {
        __builtin_cpu_init();
        unsigned a, b, c, d;
        __asm__ volatile("cpuid" : "=a"(a), "=b"(b), "=c"(c), "=d"(d) : "a"(7), "c"(0));
        bool sha = (b >> 29) & 1;
        v = sha && __builtin_cpu_supports("sse4.1") && __builtin_cpu_supports("ssse3");
        if (getenv("HZ_NO_SHANI")) v = 0;
    }
    return v == 1;
}

void sha256_host(const uint8_t* data, size_t n, uint8_t out[32]) {
    uint32_t h[8] = {0x6a09e667,0xbb67ae85,0x3c6ef372,0xa54ff53a,0x510e527f,0x9b05688c,0x1f83d9ab,0x5be0cd19};
    auto blocks = have_sha_ni() ? sha_blocks_ni : sha_blocks_portable;
    size_t full = n / 64;
    if (full) blocks(h, data, full);
    uint8_t tail[128];
    size_t rem = n - full * 64;
    memcpy(tail, data + full * 64, rem);
    tail[rem] = 0x80;
    size_t tl = rem + 9 <= 64 ? 64 : 128;
    memset(tail + rem + 1, 0, tl - rem - 1);
    uint64_t bits = (uint64_t)n * 8;
    for (int i = 0; i < 8; ++i) tail[tl - 1 - i] = (uint8_t)(bits >> (8 * i));
    blocks(h, tail, tl / 64);
    for (int i = 0; i < 8; ++i) { out[4*i] = h[i] >> 24; out[4*i+1] = h[i] >> 16; out[4*i+2] = h[i] >> 8; out[4*i+3] = h[i]; }
}

// hash `count` chunks of a buffer in parallel
void sha256_chunks_host(const uint8_t* base, const uint64_t* off, const uint32_t* size, size_t count, uint8_t* digests) {
    unsigned hw = std::thread::hardware_concurrency();
    size_t T = std::min<size_t>(std::max(1u, hw), std::min<size_t>(count, 32));
    if (T <= 1) { for (size_t i = 0; i < count; ++i) sha256_host(base + off[i], size[i], digests + 32 * i); return; }
    std::atomic<size_t> next(0);
    std::vector<std::thread> pool;
    for (size_t t = 0; t < T; ++t) pool.emplace_back([&]() {
        for (;;) { size_t i = next.fetch_add(1); if (i >= count) break; sha256_host(base + off[i], size[i], digests + 32 * i); }
    });
    for (auto& th : pool) th.join();
}

// ================================================================================================
// Container
// ================================================================================================
struct ChunkMeta {
    uint32_t index; uint64_t origOff; uint32_t origSize; uint64_t compOff; uint32_t compSize;
    uint8_t sha[32]; uint8_t len[256];
};
struct Header {
    std::string name; uint64_t size = 0; int64_t mtime = 0; uint32_t chunk = 0; uint8_t global[32];
    std::vector<ChunkMeta> chunks;
};

void put32(std::vector<uint8_t>& v, uint32_t x) { for (int i = 3; i >= 0; --i) v.push_back((uint8_t)(x >> (8 * i))); }
void put64(std::vector<uint8_t>& v, uint64_t x) { for (int i = 7; i >= 0; --i) v.push_back((uint8_t)(x >> (8 * i))); }

// CompressionHeader.writeTo (core/CompressionHeader.java:51-85) + the trailing footer pointer
// (service/cpu/CpuCompressionService.java:166-174)
void write_footer(const Header& h, uint64_t footer_pos, std::vector<uint8_t>& f) {
    f.reserve(f.size() + 76 + h.name.size() + 572 * h.chunks.size());
    put32(f, 0x44435A46u); put32(f, 1);
    put32(f, (uint32_t)h.name.size()); f.insert(f.end(), h.name.begin(), h.name.end());
    put64(f, h.size); put64(f, (uint64_t)h.mtime); put32(f, h.chunk);
    f.insert(f.end(), h.global, h.global + 32);
    put32(f, (uint32_t)h.chunks.size());
    for (const ChunkMeta& c : h.chunks) {
        put32(f, c.index); put64(f, c.origOff); put32(f, c.origSize); put64(f, c.compOff); put32(f, c.compSize);
        f.insert(f.end(), c.sha, c.sha + 32);
        for (int s = 0; s < 256; ++s) { f.push_back(0); f.push_back(c.len[s]); }       // writeShort(len)
    }
    put64(f, footer_pos);
}

struct Rd {
    const uint8_t* p; size_t n; size_t pos = 0; bool eof = false;
    uint64_t get(int bytes) {
        if (pos + bytes > n) { eof = true; pos = n; return 0; }
        uint64_t x = 0;
        for (int i = 0; i < bytes; ++i) x = x << 8 | p[pos++];
        return x;
    }
};

// CompressionHeader.readFrom (core/CompressionHeader.java:90-144).
// 0 ok; HZ_ERR_FORMAT with a reason otherwise (EOF == the reference's EOFException).
int read_header(Rd& r, Header& h, std::string& why) {
    if ((uint32_t)r.get(4) != 0x44435A46u || r.eof) { why = "Invalid file format: bad magic number"; return HZ_ERR_FORMAT; }
    uint32_t ver = (uint32_t)r.get(4);
    if (r.eof || ver != 1) { why = "Unsupported version: " + std::to_string(ver); return HZ_ERR_FORMAT; }
    int32_t nl = (int32_t)r.get(4);
    if (r.eof || nl < 0 || r.pos + (size_t)nl > r.n) { why = "truncated header"; return HZ_ERR_FORMAT; }
    h.name.assign((const char*)r.p + r.pos, (size_t)nl); r.pos += (size_t)nl;
    h.size = r.get(8); h.mtime = (int64_t)r.get(8); h.chunk = (uint32_t)r.get(4);
    if (r.pos + 32 > r.n) { why = "truncated header"; return HZ_ERR_FORMAT; }
    memcpy(h.global, r.p + r.pos, 32); r.pos += 32;
    int32_t K = (int32_t)r.get(4);
    if (r.eof || K < 0) { why = "truncated header"; return HZ_ERR_FORMAT; }
    if ((uint64_t)K * 572 > r.n - r.pos) { why = "truncated header"; return HZ_ERR_FORMAT; }
    h.chunks.resize((size_t)K);
    for (int32_t i = 0; i < K; ++i) {
        ChunkMeta& c = h.chunks[(size_t)i];
        c.index = (uint32_t)r.get(4); c.origOff = r.get(8); c.origSize = (uint32_t)r.get(4);
        c.compOff = r.get(8); c.compSize = (uint32_t)r.get(4);
        memcpy(c.sha, r.p + r.pos, 32); r.pos += 32;
        for (int s = 0; s < 256; ++s) {
            int16_t v = (int16_t)r.get(2);
            c.len[s] = (v < 0 || v > 255) ? 255 : (uint8_t)v;      // >32 is rejected by the decoder
        }
    }
    if (r.eof) { why = "truncated header"; return HZ_ERR_FORMAT; }
    return HZ_OK;
}

// ---- byte sources ----------------------------------------------------------------------------
// Large transfers are cut into slices handled by a few threads: one thread copying out of / into the page
// cache moves 1.5-4 GB/s, which is what used to bound the file pipeline.
static const uint64_t IO_SLICE_MIN = 8ull << 20;
template <class F>
bool io_parallel(uint64_t n, F&& piece) {              // piece(offset, bytes) -> bool, on up to 8 threads
    unsigned hw = std::thread::hardware_concurrency();
    uint64_t T = std::min<uint64_t>(std::min<uint64_t>(hw ? hw : 4, 8), (n + IO_SLICE_MIN - 1) / IO_SLICE_MIN);
    if (T <= 1) return piece(0, n);
    const uint64_t per = ((n + T - 1) / T + 4095) & ~4095ull;
    std::atomic<bool> ok{true};
    std::vector<std::thread> pool;
    for (uint64_t i = 0; i < T; ++i) {
        const uint64_t lo = i * per;
        if (lo >= n) break;
        const uint64_t len = std::min<uint64_t>(per, n - lo);
        pool.emplace_back([&, lo, len] { if (!piece(lo, len)) ok = false; });
    }
    for (auto& th : pool) th.join();
    return ok;
}

struct Source {                     // random-access reader over a file or a memory buffer
    int fd = -1; const uint8_t* mem = nullptr; uint64_t size = 0;
    bool read_serial(uint64_t off, void* dst, uint64_t n) const {
        if (mem) { memcpy(dst, mem + off, n); return true; }
        uint8_t* d = (uint8_t*)dst;
        while (n) {
            ssize_t r = pread(fd, d, n > (1u << 30) ? (1u << 30) : n, (off_t)off);
            if (r <= 0) { if (r < 0 && errno == EINTR) continue; return false; }
            d += r; off += (uint64_t)r; n -= (uint64_t)r;
        }
        return true;
    }
    bool read(uint64_t off, void* dst, uint64_t n) const {
        if (off > size || n > size - off) return false;            // no wrap-around: off comes from an untrusted footer
        return io_parallel(n, [&](uint64_t lo, uint64_t len) { return read_serial(off + lo, (uint8_t*)dst + lo, len); });
    }
};
struct Sink {                       // in-order writer to a file or a growing memory buffer
    int fd = -1; std::vector<uint8_t>* mem = nullptr; uint64_t pos = 0;
    bool write_serial(uint64_t off, const uint8_t* s, uint64_t n) {
        while (n) {
            ssize_t w = ::pwrite(fd, s, n > (1u << 30) ? (1u << 30) : n, (off_t)off);
            if (w <= 0) { if (w < 0 && errno == EINTR) continue; return false; }
            s += w; off += (uint64_t)w; n -= (uint64_t)w;
        }
        return true;
    }
    bool write(const void* src, uint64_t n) {
        const uint64_t at = pos;
        pos += n;
        if (mem) {
            mem->resize((size_t)(at + n));
            uint8_t* d = mem->data() + at;
            return io_parallel(n, [&](uint64_t lo, uint64_t len) { memcpy(d + lo, (const uint8_t*)src + lo, len); return true; });
        }
        return io_parallel(n, [&](uint64_t lo, uint64_t len) { return write_serial(at + lo, (const uint8_t*)src + lo, len); });
    }
};

int pin_reserve(hz_ctx* ctx, size_t bytes) {
    if (ctx->h_pin_cap >= bytes) return HZ_OK;
    if (ctx->h_pin) { cudaFreeHost(ctx->h_pin); ctx->h_pin = nullptr; ctx->h_pin_cap = 0; }
    cudaError_t e = cudaHostAlloc(&ctx->h_pin, bytes, cudaHostAllocDefault);
    if (e != cudaSuccess) return hz_cuda_fail(ctx, e, "cudaHostAlloc(staging)");
    ctx->h_pin_cap = bytes;
    return HZ_OK;
}

// ---- stage metrics (model/StageMetrics.java) -----------------------------------------------------
// Host stages are wall-clock times of the pipeline's tasks (reader / writer threads add theirs atomically);
// kernel stages come from the library's event pairs, collected for the duration of a file-level call.
struct HostClock {
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    double ms() const { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); }
};
struct IoAcc { std::atomic<uint64_t> us{0}, bytes{0}, count{0}; void add(double ms, uint64_t b) { us += (uint64_t)(ms * 1e3); bytes += b; ++count; } };

int stage_of_kernel(const char* name, bool decompressing) {
    if (!strncmp(name, "hist", 4) || !strcmp(name, "sum_seg_hist") || !strncmp(name, "global_", 7) || !strncmp(name, "nccl", 4)) return HZ_STAGE_FREQUENCY_ANALYSIS;
    if (!strncmp(name, "codebook", 8) || !strcmp(name, "chunk_offsets") || !strcmp(name, "codes_from_lengths")) return HZ_STAGE_HUFFMAN_TREE_BUILD;
    if (!strcmp(name, "dec_tables") || !strcmp(name, "dec_plan") || !strcmp(name, "dec_ident") || !strcmp(name, "dec_zero")) return HZ_STAGE_HUFFMAN_TREE_BUILD;   // :518
    if (!strcmp(name, "encode")) return HZ_STAGE_ENCODING;
    if (!strncmp(name, "dec_", 4)) return HZ_STAGE_DECODING;
    if (!strncmp(name, "sha256", 6)) return decompressing ? HZ_STAGE_CHECKSUM_VERIFY : HZ_STAGE_CHECKSUM_COMPUTE;
    return -1;
}

struct StageScope {                 // collects the kernels' event pairs of one file-level call into ctx->stages
    hz_ctx* ctx; bool was; std::vector<hz_prof_entry> saved; uint64_t bytes; bool decompressing;
    StageScope(hz_ctx* c, uint64_t nbytes, bool dec = false) : ctx(c), was(c->prof), bytes(nbytes), decompressing(dec) {
        hz_prof_resolve(c);
        saved.swap(c->prof_entries);
        c->prof = true;
        c->stages = StageAcc();
    }
    ~StageScope() {
        hz_prof_resolve(ctx);
        for (const auto& e : ctx->prof_entries) {
            const int s = stage_of_kernel(e.name, decompressing);
            if (s >= 0) { ctx->stages.ms[s] += e.ms; ctx->stages.count[s] += e.launches; ctx->stages.bytes[s] = bytes; }
        }
        if (was) for (const auto& e : ctx->prof_entries) {          // a caller that profiles keeps seeing these launches
            bool found = false;
            for (auto& o : saved) if (!strcmp(o.name, e.name)) { o.ms += e.ms; o.launches += e.launches; found = true; break; }
            if (!found) saved.push_back(e);
        }
        ctx->prof_entries.swap(saved);
        ctx->prof = was;
    }
};

// SHA-256 of a batch's chunks: on the GPU (hz_sha256.cu, one thread per chunk, ~27 MB/s per chunk) when the batch has
// enough chunks to beat the host's SHA units (~1.8 GB/s per core, one chunk per thread), else on the host while the
// GPU codes the batch.  Measured on a B200 box with 16 host cores (profiles/r02_sha256_gpu_vs_host.txt, 1 GiB):
// 16 KiB chunks 700 vs 29 GB/s, 64 KiB 443 vs 29, 256 KiB 108 vs 29, 1 MiB 27 vs 29, 16 MiB 1.8 vs 29: the kernel wins
// from about a thousand chunks per call, i.e. chunks <= 64 KiB with 128 MiB batches.
const size_t kShaGpuMinChunks = 1536;

// bytes of input handled per GPU batch (>= one chunk)
uint64_t batch_bytes_for(uint32_t chunk_bytes) {
    const uint64_t target = 128ull << 20;      // two pinned slots of this size are in flight
    if (chunk_bytes >= target) return chunk_bytes;
    return (target / chunk_bytes) * chunk_bytes;
}

// ---- compress --------------------------------------------------------------------------------
int compress_core(hz_ctx* ctx, const Source& src, Sink& dst, uint32_t chunk_bytes, const std::string& name,
                  int64_t mtime_ms, hz_progress_fn progress, void* user) {
    if (chunk_bytes == 0 || chunk_bytes > 0x7fffffffu) return hz_fail(ctx, HZ_ERR_ARG, "bad chunk size %u", chunk_bytes);
    HZ_CUDA(ctx, cudaSetDevice(ctx->device));
    const uint64_t n = src.size;
    const uint64_t K = hz_num_chunks(n, chunk_bytes);                        // :64
    if (K > 0x7fffffffull) return hz_fail(ctx, HZ_ERR_ARG, "too many chunks");
    Header h; h.name = name; h.size = n; h.mtime = mtime_ms; h.chunk = chunk_bytes;
    h.chunks.resize((size_t)K);
    const uint64_t bb = batch_bytes_for(chunk_bytes);
    const uint64_t max_batch = std::min<uint64_t>(bb, n);
    const size_t kb_max = (size_t)(bb / chunk_bytes);
    // Two pinned slots, each [input batch][payload batch][comp_off][len]: while the GPU encodes batch b and the
    // host hashes it, a reader task fills the other slot with batch b+1 and a writer task drains batch b-1.
    const size_t off_in = 0, off_out = (size_t)((max_batch + 255) & ~255ull);
    const size_t off_coff = off_out + (size_t)((max_batch + 16 + 255) & ~255ull);
    const size_t off_len = off_coff + (((kb_max + 1) * 8 + 255) & ~(size_t)255);
    const size_t slot_bytes = (off_len + kb_max * 256 + 256 + 4095) & ~(size_t)4095;
    const uint64_t nbatch = (n + bb - 1) / bb;
    HZ_TRY(pin_reserve(ctx, slot_bytes * (nbatch > 1 ? 2 : 1)));
    uint8_t* pin_base = (uint8_t*)ctx->h_pin;
    DevBuf& d_in = ctx->stage_in; DevBuf& d_out = ctx->stage_out;
    DevBuf& d_off = ctx->stage_d; DevBuf& d_len = ctx->stage_e;
    HZ_TRY(hz_reserve(ctx, &d_in, max_batch));
    HZ_TRY(hz_reserve(ctx, &d_out, max_batch + 16));
    HZ_TRY(hz_reserve(ctx, &d_off, (kb_max + 1) * 8));
    HZ_TRY(hz_reserve(ctx, &d_len, kb_max * 256));

    uint64_t comp_total = 0, done = 0;
    IoAcc io;                                         // declared before the futures: the tasks reference it
    StageScope scope(ctx, n);
    std::future<bool> fut_rd, fut_wr;                 // std::async futures join in their destructor on every exit path
    if (n) fut_rd = std::async(std::launch::async, [&src, &io, pin_base, off_in, bb, n] {
        HostClock c; const bool ok = src.read(0, pin_base + off_in, std::min<uint64_t>(bb, n)); io.add(c.ms(), std::min<uint64_t>(bb, n)); return ok; });
    uint64_t b = 0;
    for (uint64_t pos = 0, k0 = 0; pos < n; ++b) {
        uint8_t* pin = pin_base + (b & 1) * slot_bytes;
        const uint64_t bn = std::min<uint64_t>(bb, n - pos);
        const size_t kb = (size_t)hz_num_chunks(bn, chunk_bytes);
        if (!fut_rd.get()) return hz_fail(ctx, HZ_ERR_IO, "read failed at offset %llu", (unsigned long long)pos);
        if (pos + bn < n) {                           // the other slot's input was consumed (copied, hashed) a batch ago
            uint8_t* nxt = pin_base + ((b + 1) & 1) * slot_bytes + off_in;
            const uint64_t npos = pos + bn, nn = std::min<uint64_t>(bb, n - npos);
            fut_rd = std::async(std::launch::async, [&src, &io, nxt, npos, nn] {
                HostClock c; const bool ok = src.read(npos, nxt, nn); io.add(c.ms(), nn); return ok; });
        }
        HZ_CUDA(ctx, cudaMemcpyAsync(d_in.p, pin + off_in, bn, cudaMemcpyHostToDevice, ctx->stream));
        HZ_TRY(hz_encode(ctx, (const uint8_t*)d_in.p, bn, chunk_bytes, (uint8_t*)d_out.p, bn + 16,
                         (uint64_t*)d_off.p, (uint8_t*)d_len.p, nullptr));
        HZ_CUDA(ctx, cudaMemcpyAsync(pin + off_coff, d_off.p, (kb + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream));
        HZ_CUDA(ctx, cudaMemcpyAsync(pin + off_len, d_len.p, kb * 256, cudaMemcpyDeviceToHost, ctx->stream));
        // SHA-256 of the plaintext chunks (:226-228): on the host while the GPU encodes, or — thousands of small
        // chunks per batch — by the GPU kernel on the batch that is already in device memory
        std::vector<uint64_t> so(kb); std::vector<uint32_t> ss(kb); std::vector<uint8_t> dig(kb * 32);
        for (size_t i = 0; i < kb; ++i) { so[i] = (uint64_t)i * chunk_bytes; ss[i] = (uint32_t)std::min<uint64_t>(chunk_bytes, bn - so[i]); }
        if (kb >= kShaGpuMinChunks) {
            HZ_TRY(hz_reserve(ctx, &ctx->stage_a, kb * 32));
            HZ_TRY(hzk_sha256(ctx, (const uint8_t*)d_in.p, bn, chunk_bytes, (uint32_t)kb, (uint8_t*)ctx->stage_a.p));
            HZ_CUDA(ctx, cudaMemcpyAsync(dig.data(), ctx->stage_a.p, kb * 32, cudaMemcpyDeviceToHost, ctx->stream));
        } else {
            HostClock c; sha256_chunks_host(pin + off_in, so.data(), ss.data(), kb, dig.data()); ctx->stages.add(HZ_STAGE_CHECKSUM_COMPUTE, c.ms(), bn);
        }
        HZ_TRY(hz_sync(ctx));
        const uint64_t* coff = (const uint64_t*)(pin + off_coff);
        const uint64_t btotal = coff[kb];
        // this slot's payload region was handed to the writer two batches ago; that write was joined last batch
        HZ_CUDA(ctx, cudaMemcpyAsync(pin + off_out, d_out.p, btotal, cudaMemcpyDeviceToHost, ctx->stream));
        HZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (fut_wr.valid() && !fut_wr.get()) return hz_fail(ctx, HZ_ERR_IO, "write failed");
        fut_wr = std::async(std::launch::async, [&dst, &io, pin, off_out, btotal] {
            HostClock c; const bool ok = dst.write(pin + off_out, btotal); io.add(c.ms(), btotal); return ok; });
        for (size_t i = 0; i < kb; ++i) {
            ChunkMeta& c = h.chunks[(size_t)k0 + i];
            c.index = (uint32_t)(k0 + i);
            c.origOff = (k0 + i) * (uint64_t)chunk_bytes;
            c.origSize = ss[i];
            c.compOff = comp_total + coff[i];
            c.compSize = (uint32_t)(coff[i + 1] - coff[i]);
            memcpy(c.sha, &dig[32 * i], 32);
            memcpy(c.len, pin + off_len + 256 * i, 256);
            if (progress) progress((double)(++done) / (double)K, user);      // :111-114
        }
        comp_total += btotal; pos += bn; k0 += kb;
    }
    if (fut_wr.valid() && !fut_wr.get()) return hz_fail(ctx, HZ_ERR_IO, "write failed");
    // global checksum = SHA-256 over the chunk digests in index order (:106-109,126)
    std::vector<uint8_t> cat(32 * (size_t)K);
    for (size_t i = 0; i < (size_t)K; ++i) memcpy(&cat[32 * i], h.chunks[i].sha, 32);
    sha256_host(cat.data(), cat.size(), h.global);
    HostClock hc;
    std::vector<uint8_t> footer;
    write_footer(h, comp_total, footer);
    if (!dst.write(footer.data(), footer.size())) return hz_fail(ctx, HZ_ERR_IO, "write failed (footer)");
    ctx->stages.add(HZ_STAGE_HEADER_WRITE, hc.ms(), footer.size());                 // :178
    ctx->stages.ms[HZ_STAGE_FILE_IO] += io.us / 1e3; ctx->stages.count[HZ_STAGE_FILE_IO] += io.count; ctx->stages.bytes[HZ_STAGE_FILE_IO] += io.bytes;
    return HZ_OK;
}

// ---- decompress --------------------------------------------------------------------------------
// Container probe exactly as CpuCompressionService.decompress (:337-393).
int parse_container(hz_ctx* ctx, const Source& src, Header& h, uint64_t* data_start) {
    const uint64_t n = src.size;
    std::string why;
    {
        size_t bl = (size_t)std::min<uint64_t>(n, 64 * 1024);
        std::vector<uint8_t> buf(bl, 0);
        size_t rl = std::min<size_t>(bl, 4096);
        if (rl && !src.read(0, buf.data(), rl)) return hz_fail(ctx, HZ_ERR_IO, "read failed");
        Rd r{buf.data(), bl};
        Header t;
        if (read_header(r, t, why) == HZ_OK) {                                 // header-first (legacy)
            uint64_t tot = 0;
            for (auto& c : t.chunks) tot += c.compSize;
            if (tot > n) return hz_fail(ctx, HZ_ERR_FORMAT, "Invalid file format: chunk table exceeds file size");
            *data_start = n - tot;
            h = std::move(t);
            return HZ_OK;
        }
    }
    if (n < 8) return hz_fail(ctx, HZ_ERR_FORMAT, "Invalid file format: file too small");
    uint8_t p8[8];
    if (!src.read(n - 8, p8, 8)) return hz_fail(ctx, HZ_ERR_IO, "read failed");
    int64_t fs = 0;
    for (int i = 0; i < 8; ++i) fs = (int64_t)(((uint64_t)fs << 8) | p8[i]);
    if (fs < 0 || (uint64_t)fs >= n - 8)                                        // :372-374
        return hz_fail(ctx, HZ_ERR_FORMAT, "Invalid footer position: %lld", (long long)fs);
    std::vector<uint8_t> fb((size_t)(n - 8 - (uint64_t)fs));
    if (!src.read((uint64_t)fs, fb.data(), fb.size())) return hz_fail(ctx, HZ_ERR_IO, "read failed");
    Rd r{fb.data(), fb.size()};
    if (read_header(r, h, why) != HZ_OK) return hz_fail(ctx, HZ_ERR_FORMAT, "%s", why.c_str());
    *data_start = 0;
    return HZ_OK;
}

int decompress_core(hz_ctx* ctx, const Source& src, Sink* dst, hz_progress_fn progress, void* user) {
    HZ_CUDA(ctx, cudaSetDevice(ctx->device));
    Header h; uint64_t data_start = 0;
    IoAcc io;                                         // declared before the futures: the tasks reference it
    StageScope scope(ctx, 0, true);
    { HostClock c; const int prc = parse_container(ctx, src, h, &data_start); ctx->stages.add(HZ_STAGE_FILE_IO, c.ms(), 0); HZ_TRY(prc); }   // :395
    scope.bytes = h.size;
    const size_t K = h.chunks.size();
    // batches of consecutive chunks, up to ~128 MiB of output each; two pinned slots [compressed][output]:
    // a reader task fetches batch b+1 while batch b is decoded and verified, a writer task drains batch b-1
    struct Batch { size_t k0, k1; uint64_t ob, cb; };
    std::vector<Batch> batches;
    uint64_t max_ob = 0, max_cb = 0;
    for (size_t k0 = 0; k0 < K; ) {
        size_t k1 = k0; uint64_t ob = 0, cb = 0;
        while (k1 < K && (k1 == k0 || ob + h.chunks[k1].origSize <= (128ull << 20))) {
            ob += h.chunks[k1].origSize; cb += h.chunks[k1].compSize; ++k1;
        }
        batches.push_back({k0, k1, ob, cb});
        max_ob = std::max(max_ob, ob); max_cb = std::max(max_cb, cb);
        k0 = k1;
    }
    const size_t o_comp = 0, o_out = (size_t)((max_cb + 255) & ~255ull);
    const size_t slot_bytes = (o_out + max_ob + 256 + 4095) & ~(size_t)4095;
    if (!batches.empty()) HZ_TRY(pin_reserve(ctx, slot_bytes * (batches.size() > 1 ? 2 : 1)));
    uint8_t* pin_base = (uint8_t*)ctx->h_pin;
    // the chunks of a batch are gathered back to back; consecutive chunks of a well-formed file are one extent
    auto read_batch = [&src, &h, &io, data_start](const Batch& B, uint8_t* dstp) -> long {
        HostClock clk;
        uint64_t ca = 0;
        size_t i = B.k0;
        while (i < B.k1) {
            size_t j = i; uint64_t ext = 0;
            const uint64_t start = h.chunks[i].compOff;
            while (j < B.k1 && h.chunks[j].compOff == start + ext) { ext += h.chunks[j].compSize; ++j; }
            if (start > UINT64_MAX - data_start || !src.read(data_start + start, dstp + ca, ext)) return (long)i;   // :429-436
            ca += ext; i = j;
        }
        io.add(clk.ms(), ca);
        return -1;
    };
    std::future<long> fut_rd;
    std::future<bool> fut_wr;
    if (!batches.empty()) fut_rd = std::async(std::launch::async, read_batch, batches[0], pin_base + o_comp);
    size_t done = 0;
    for (size_t b = 0; b < batches.size(); ++b) {
        const Batch& B = batches[b];
        const size_t k0 = B.k0, k1 = B.k1, kb = k1 - k0;
        uint8_t* pin = pin_base + (b & 1) * slot_bytes;
        const long bad = fut_rd.get();
        if (bad >= 0) return hz_fail(ctx, HZ_ERR_IO, "chunk %ld: compressed data out of file bounds", bad);
        if (b + 1 < batches.size())
            fut_rd = std::async(std::launch::async, read_batch, batches[b + 1], pin_base + ((b + 1) & 1) * slot_bytes + o_comp);
        std::vector<uint64_t> coff(kb), ooff(kb); std::vector<uint32_t> csz(kb), osz(kb); std::vector<uint8_t> lens(kb * 256);
        uint64_t ca = 0, oa = 0;
        for (size_t i = 0; i < kb; ++i) {
            const ChunkMeta& c = h.chunks[k0 + i];
            coff[i] = ca; csz[i] = c.compSize; ooff[i] = oa; osz[i] = c.origSize;
            memcpy(&lens[256 * i], c.len, 256);
            ca += c.compSize; oa += c.origSize;
        }
        // this slot's output region was handed to the writer two batches ago; that write was joined last batch
        // thousands of equal-sized chunks: decode into device memory and hash them there (hz_sha256.cu), else through
        // the host-buffer call (pipelined copies) and the host's SHA units
        bool gpu_sha = kb >= kShaGpuMinChunks;
        for (size_t i = 0; gpu_sha && i + 1 < kb; ++i) gpu_sha = osz[i] == osz[0];
        if (gpu_sha && osz[kb - 1] > osz[0]) gpu_sha = false;
        std::vector<uint8_t> dig(kb * 32);
        int rc;
        if (gpu_sha) {
            rc = hz_reserve(ctx, &ctx->pipe_in[0], B.cb + 32);
            if (rc == HZ_OK) rc = hz_reserve(ctx, &ctx->pipe_out[0], B.ob + 16);
            if (rc == HZ_OK) rc = hz_reserve(ctx, &ctx->pipe_meta_d, kb * 32);
            if (rc != HZ_OK) return rc;
            HZ_CUDA(ctx, cudaMemcpyAsync(ctx->pipe_in[0].p, pin + o_comp, B.cb, cudaMemcpyHostToDevice, ctx->stream));
            rc = hz_decode(ctx, (const uint8_t*)ctx->pipe_in[0].p, B.cb, coff.data(), csz.data(), osz.data(), ooff.data(), lens.data(),
                           (uint32_t)kb, (uint8_t*)ctx->pipe_out[0].p, B.ob);
            if (rc == HZ_OK) rc = hzk_sha256(ctx, (const uint8_t*)ctx->pipe_out[0].p, B.ob, osz[0], (uint32_t)kb, (uint8_t*)ctx->pipe_meta_d.p);
            if (rc == HZ_OK) {
                HZ_CUDA(ctx, cudaMemcpyAsync(dig.data(), ctx->pipe_meta_d.p, kb * 32, cudaMemcpyDeviceToHost, ctx->stream));
                HZ_CUDA(ctx, cudaMemcpyAsync(pin + o_out, ctx->pipe_out[0].p, B.ob, cudaMemcpyDeviceToHost, ctx->stream));
                rc = hz_sync(ctx);
            }
        } else {
            rc = hz_decode(ctx, pin + o_comp, B.cb, coff.data(), csz.data(), osz.data(), ooff.data(), lens.data(),
                           (uint32_t)kb, pin + o_out, B.ob);
        }
        if (rc != HZ_OK) {
            if (rc == HZ_ERR_DECODE || rc == HZ_ERR_BAD_LENGTHS)
                return hz_fail(ctx, rc, "Chunk decompression failed: %s (chunks %zu..%zu)", hz_strerror(rc), k0, k1 - 1);
            return rc;
        }
        if (!gpu_sha) { HostClock c; sha256_chunks_host(pin + o_out, ooff.data(), osz.data(), kb, dig.data()); ctx->stages.add(HZ_STAGE_CHECKSUM_VERIFY, c.ms(), B.ob); }   // :536-550
        for (size_t i = 0; i < kb; ++i)
            if (memcmp(&dig[32 * i], h.chunks[k0 + i].sha, 32) != 0)
                return hz_fail(ctx, HZ_ERR_CHECKSUM, "Checksum mismatch in chunk %zu", k0 + i);
        if (dst) {
            if (fut_wr.valid() && !fut_wr.get()) return hz_fail(ctx, HZ_ERR_IO, "write failed");
            const uint64_t ob = B.ob;
            fut_wr = std::async(std::launch::async, [dst, &io, pin, o_out, ob] {
                HostClock c; const bool ok = dst->write(pin + o_out, ob); io.add(c.ms(), ob); return ok; });
        }
        for (size_t i = 0; i < kb; ++i)
            if (progress) progress((double)(++done) / (double)K, user);                  // :464-467
    }
    if (fut_wr.valid() && !fut_wr.get()) return hz_fail(ctx, HZ_ERR_IO, "write failed");
    ctx->stages.ms[HZ_STAGE_FILE_IO] += io.us / 1e3; ctx->stages.count[HZ_STAGE_FILE_IO] += io.count; ctx->stages.bytes[HZ_STAGE_FILE_IO] += io.bytes;
    return HZ_OK;
}

std::string base_name(const char* path) {
    std::string s(path);
    size_t p = s.find_last_of('/');
    return p == std::string::npos ? s : s.substr(p + 1);
}

}  // namespace

extern "C" {

int hz_compress_file(hz_ctx* ctx, const char* in_path, const char* out_path, uint32_t chunk_bytes,
                     const char* name_override, int64_t mtime_ms_override, hz_progress_fn progress, void* user) {
    if (!ctx || !in_path || !out_path) return hz_fail(ctx, HZ_ERR_ARG, "hz_compress_file: bad argument");
    Source src; src.fd = open(in_path, O_RDONLY);
    if (src.fd < 0) return hz_fail(ctx, HZ_ERR_IO, "cannot open %s: %s", in_path, strerror(errno));
    struct stat st;
    if (fstat(src.fd, &st) != 0) { close(src.fd); return hz_fail(ctx, HZ_ERR_IO, "cannot stat %s", in_path); }
    src.size = (uint64_t)st.st_size;
    int64_t mtime = mtime_ms_override >= 0 ? mtime_ms_override
                                           : (int64_t)st.st_mtim.tv_sec * 1000 + st.st_mtim.tv_nsec / 1000000;   // :73
    Sink dst; dst.fd = open(out_path, O_WRONLY | O_CREAT | O_TRUNC, 0644);
    if (dst.fd < 0) { close(src.fd); return hz_fail(ctx, HZ_ERR_IO, "cannot create %s: %s", out_path, strerror(errno)); }
    int rc = compress_core(ctx, src, dst, chunk_bytes, name_override ? std::string(name_override) : base_name(in_path),
                           mtime, progress, user);
    close(src.fd);
    if (close(dst.fd) != 0 && rc == HZ_OK) rc = hz_fail(ctx, HZ_ERR_IO, "close failed");
    return rc;
}

int hz_decompress_file(hz_ctx* ctx, const char* in_path, const char* out_path, hz_progress_fn progress, void* user) {
    if (!ctx || !in_path || !out_path) return hz_fail(ctx, HZ_ERR_ARG, "hz_decompress_file: bad argument");
    Source src; src.fd = open(in_path, O_RDONLY);
    if (src.fd < 0) return hz_fail(ctx, HZ_ERR_IO, "cannot open %s: %s", in_path, strerror(errno));
    struct stat st; fstat(src.fd, &st); src.size = (uint64_t)st.st_size;
    Sink dst; dst.fd = open(out_path, O_WRONLY | O_CREAT | O_TRUNC, 0644);
    if (dst.fd < 0) { close(src.fd); return hz_fail(ctx, HZ_ERR_IO, "cannot create %s: %s", out_path, strerror(errno)); }
    int rc = decompress_core(ctx, src, &dst, progress, user);
    close(src.fd);
    if (close(dst.fd) != 0 && rc == HZ_OK) rc = hz_fail(ctx, HZ_ERR_IO, "close failed");
    return rc;
}

int hz_verify_file(hz_ctx* ctx, const char* path, int* ok) {
    if (!ctx || !path || !ok) return hz_fail(ctx, HZ_ERR_ARG, "hz_verify_file: bad argument");
    *ok = 0;
    Source src; src.fd = open(path, O_RDONLY);
    if (src.fd < 0) return hz_fail(ctx, HZ_ERR_IO, "cannot open %s: %s", path, strerror(errno));
    struct stat st; fstat(src.fd, &st); src.size = (uint64_t)st.st_size;
    int rc = decompress_core(ctx, src, nullptr, nullptr, nullptr);
    close(src.fd);
    if (rc == HZ_OK) { *ok = 1; return HZ_OK; }
    if (rc == HZ_ERR_CHECKSUM || rc == HZ_ERR_DECODE || rc == HZ_ERR_BAD_LENGTHS || rc == HZ_ERR_FORMAT) return HZ_OK;
    return rc;
}

int hz_compress_buffer(hz_ctx* ctx, const uint8_t* data, uint64_t n, uint32_t chunk_bytes, const char* name,
                       int64_t mtime_ms, uint8_t** out, uint64_t* out_n) {
    if (!ctx || !out || !out_n || (n && !data) || !name) return hz_fail(ctx, HZ_ERR_ARG, "hz_compress_buffer: bad argument");
    Source src; src.mem = data; src.size = n;
    std::vector<uint8_t> v;
    Sink dst; dst.mem = &v;
    HZ_TRY(compress_core(ctx, src, dst, chunk_bytes, name, mtime_ms, nullptr, nullptr));
    *out = (uint8_t*)malloc(v.size() ? v.size() : 1);
    if (!*out) return hz_fail(ctx, HZ_ERR_NOMEM, "malloc failed");
    memcpy(*out, v.data(), v.size());
    *out_n = v.size();
    return HZ_OK;
}

int hz_decompress_buffer(hz_ctx* ctx, const uint8_t* dcz, uint64_t n, uint8_t** out, uint64_t* out_n) {
    if (!ctx || !out || !out_n || (n && !dcz)) return hz_fail(ctx, HZ_ERR_ARG, "hz_decompress_buffer: bad argument");
    Source src; src.mem = dcz; src.size = n;
    std::vector<uint8_t> v;
    Sink dst; dst.mem = &v;
    HZ_TRY(decompress_core(ctx, src, &dst, nullptr, nullptr));
    *out = (uint8_t*)malloc(v.size() ? v.size() : 1);
    if (!*out) return hz_fail(ctx, HZ_ERR_NOMEM, "malloc failed");
    memcpy(*out, v.data(), v.size());
    *out_n = v.size();
    return HZ_OK;
}

void hz_free(void* p) { free(p); }

int hz_stage_metrics(const hz_ctx* ctx, hz_stage_metric out[HZ_STAGE_COUNT]) {
    if (!ctx || !out) return HZ_ERR_ARG;
    for (int i = 0; i < HZ_STAGE_COUNT; ++i) { out[i].ms = ctx->stages.ms[i]; out[i].count = ctx->stages.count[i]; out[i].bytes = ctx->stages.bytes[i]; }
    return HZ_OK;
}

const char* hz_stage_name(int stage) {
    static const char* names[HZ_STAGE_COUNT] = {"FREQUENCY_ANALYSIS", "HUFFMAN_TREE_BUILD", "ENCODING", "CHECKSUM_COMPUTE",
                                                "FILE_IO", "HEADER_WRITE", "DECODING", "CHECKSUM_VERIFY"};
    return stage >= 0 && stage < HZ_STAGE_COUNT ? names[stage] : "";
}

// test hook: host SHA-256 (both the SHA-NI and the portable path are covered by tests)
void hz_host_sha256(const uint8_t* data, uint64_t n, uint8_t* out32) { sha256_host(data, (size_t)n, out32); }

}  // extern "C"
