// hz_hist.cu — per-segment byte histograms (stage 1 of the encode pipeline).
//
// Replaces FrequencyService.computeHistogram (service/FrequencyService.java:16,
// cpu/CpuFrequencyService.java:29-46, gpu/TornadoKernels.java:89-100 histogramTiledKernel),
// batched over every chunk of the input.  One CTA histograms one segment (HZ_SEG_BYTES of one
// chunk) and writes 256 uint32 bins; the codebook kernel sums a chunk's segments.  Keeping the
// per-segment bins lets the encoder derive every segment's exact output bit offset from
// sum(bins * code length) without a second pass over the data and without a look-back chain.
//
// Three kernels:
//  * hist_seg_lanes    — (default) one counter per (bin, lane): conflict-free shared-memory atomics,
//    distribution independent: 0.22 ms per GiB (4.8 TB/s) from 1 to 8 bits/symbol on B200.
//  * hist_seg_atomic   — (HZ_HIST=atomic) per-warp private uint32 bins updated with shared-memory atomics;
//    one column sum per bin at the end.  Measured on B200, 1 GiB: 0.18 ms for a constant stream
//    (same-address atomics are combined by the hardware), 0.19 ms at 1 bit/symbol, 0.33 ms at
//    4 bits/symbol, 0.42 ms for uniform bytes (bank conflicts between different bins).
//  * hist_seg_private  — per-THREAD private 8-bit counters in shared memory (64 KiB per CTA,
//    column t of a [64][256] word matrix, so lane == bank: conflict-free), plain
//    LDS.U8/IADD/STS.U8.  Distribution independent (0.61 ms per GiB) but ~7 instructions per
//    byte; kept as the measured alternative (HZ_HIST=private).  A thread sees at most 224 bytes
//    per segment (HZ_SEG_BYTES = 7*8192), so a counter cannot wrap.
#include "hz_common.cuh"
#include "hz_hist_lanes.cuh"

__device__ __forceinline__ void seg_geometry(uint64_t n, uint32_t chunk_bytes, uint32_t spc,
                                             uint64_t* seg_begin, uint32_t* seg_len) {
    uint32_t seg = blockIdx.x;
    uint32_t k = seg / spc, s = seg - k * spc;
    uint64_t cbeg = (uint64_t)k * chunk_bytes;
    uint64_t clen = n - cbeg < chunk_bytes ? n - cbeg : chunk_bytes;
    uint64_t sbeg = (uint64_t)s * HZ_SEG_BYTES;
    if (sbeg >= clen) { *seg_begin = 0; *seg_len = 0; return; }
    *seg_begin = cbeg + sbeg;
    uint64_t sl = clen - sbeg;
    *seg_len = sl < HZ_SEG_BYTES ? (uint32_t)sl : HZ_SEG_BYTES;
}

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(HZ_THREADS, 3)
hist_seg_private(const uint8_t* __restrict__ in, uint64_t n, uint32_t chunk_bytes, uint32_t spc,
                 uint32_t* __restrict__ seg_hist) {
    extern __shared__ __align__(16) uint32_t sm[];
    uint32_t* cnt = sm;                    // [64 rows][256 threads] words, 4 byte counters each
    uint32_t* extra = sm + 64 * 256;       // 256 uint32 bins for the unaligned head / tail bytes
    const uint32_t t = threadIdx.x;

    uint64_t sbeg; uint32_t slen;
    seg_geometry(n, chunk_bytes, spc, &sbeg, &slen);
    uint32_t* dst = seg_hist + (size_t)blockIdx.x * 256;
    if (slen == 0) { dst[t] = 0; return; }

    // zero the counters
    {
        uint4 z = make_uint4(0, 0, 0, 0);
        uint4* c4 = reinterpret_cast<uint4*>(cnt);
#pragma unroll
        for (int i = 0; i < 16; ++i) c4[t + i * 256] = z;
        extra[t] = 0;
    }
    __syncthreads();

    const uint8_t* p = in + sbeg;
    uint32_t head = (uint32_t)((16 - (reinterpret_cast<uintptr_t>(p) & 15)) & 15);
    if (head > slen) head = slen;
    uint32_t nvec = (slen - head) >> 4;
    uint32_t tail = slen - head - (nvec << 4);
    if (t < head) atomicAdd(&extra[p[t]], 1u);
    if (t < tail) atomicAdd(&extra[p[head + (nvec << 4) + t]], 1u);

    const uint4* pv = reinterpret_cast<const uint4*>(p + head);
    uint8_t* c8 = reinterpret_cast<uint8_t*>(cnt);
    const uint32_t t4 = t * 4;
    // nvec <= 3584 -> at most 14 vectors (224 bytes) per thread
    for (uint32_t i = t; i < nvec; i += HZ_THREADS) {
        uint4 v = ld_stream_u4(pv + i);
        uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int g = 0; g < 4; ++g) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                // z = sym * 0x101 ; byte address = (sym>>2)*1024 + t*4 + (sym&3)
                uint32_t z = __byte_perm(w[g], 0, 0x4400 | (j * 0x11));
                uint32_t a = (z & 0xFC03u) | t4;
                c8[a] = (uint8_t)(c8[a] + 1);
            }
        }
    }
    __syncthreads();

    // reduce the 256 columns of each row: thread t -> row t/4, quarter t%4 of the columns
    {
        uint32_t r = t >> 2, q = t & 3, lane = t & 31;
        uint32_t a = 0, b = 0;
        const uint32_t* row = cnt + r * 256 + q * 64;
#pragma unroll 8
        for (int j = 0; j < 64; ++j) {
            uint32_t w = row[(j + lane) & 63];
            a += w & 0x00FF00FFu;
            b += (w >> 8) & 0x00FF00FFu;
        }
        a += __shfl_xor_sync(0xffffffffu, a, 1);
        b += __shfl_xor_sync(0xffffffffu, b, 1);
        a += __shfl_xor_sync(0xffffffffu, a, 2);
        b += __shfl_xor_sync(0xffffffffu, b, 2);
        if (q == 0) {
            uint4 o;
            o.x = (a & 0xFFFFu) + extra[4 * r + 0];
            o.y = (b & 0xFFFFu) + extra[4 * r + 1];
            o.z = (a >> 16) + extra[4 * r + 2];
            o.w = (b >> 16) + extra[4 * r + 3];
            reinterpret_cast<uint4*>(dst)[r] = o;
        }
    }
}

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(HZ_THREADS)
hist_seg_atomic(const uint8_t* __restrict__ in, uint64_t n, uint32_t chunk_bytes, uint32_t spc,
                uint32_t* __restrict__ seg_hist) {
    __shared__ uint32_t h[HZ_THREADS / 32][256];
    const uint32_t t = threadIdx.x, wid = t >> 5;
    uint64_t sbeg; uint32_t slen;
    seg_geometry(n, chunk_bytes, spc, &sbeg, &slen);
    uint32_t* dst = seg_hist + (size_t)blockIdx.x * 256;
    if (slen == 0) { dst[t] = 0; return; }
#pragma unroll
    for (int i = 0; i < HZ_THREADS / 32; ++i) h[i][t] = 0;
    __syncthreads();
    const uint8_t* p = in + sbeg;
    uint32_t head = (uint32_t)((16 - (reinterpret_cast<uintptr_t>(p) & 15)) & 15);
    if (head > slen) head = slen;
    uint32_t nvec = (slen - head) >> 4;
    uint32_t tail = slen - head - (nvec << 4);
    uint32_t* mine = h[wid];
    if (t < head) atomicAdd(&mine[p[t]], 1u);
    if (t < tail) atomicAdd(&mine[p[head + (nvec << 4) + t]], 1u);
    const uint4* pv = reinterpret_cast<const uint4*>(p + head);
    for (uint32_t i = t; i < nvec; i += HZ_THREADS) {
        uint4 v = ld_stream_u4(pv + i);
        uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            atomicAdd(&mine[w[g] & 0xFF], 1u);
            atomicAdd(&mine[(w[g] >> 8) & 0xFF], 1u);
            atomicAdd(&mine[(w[g] >> 16) & 0xFF], 1u);
            atomicAdd(&mine[w[g] >> 24], 1u);
        }
    }
    __syncthreads();
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < HZ_THREADS / 32; ++i) s += h[i][t];
    dst[t] = s;
}

// ---------------------------------------------------------------------------------------------
// hist_seg_lanes — one uint32 counter per (bin, lane): word bin*32 + lane, so a lane always hits its
// own bank and every shared-memory atomic of a warp is ONE wavefront whatever the byte distribution
// (hist_seg_atomic needs ~2.5 for 32 random bins over 32 banks).  32 KiB per CTA, all warps of the
// CTA share the copy (atomics resolve collisions between warps).  Per symbol: PRMT (byte extract),
// IMAD (bin*128 + lane address), RED.  The 32 columns of a bin are summed with a rotated,
// conflict-free read at the end.
__global__ void __launch_bounds__(HZ_THREADS)
hist_seg_lanes(const uint8_t* __restrict__ in, uint64_t n, uint32_t chunk_bytes, uint32_t spc, uint32_t mult,
               uint32_t* __restrict__ seg_hist) {
    __shared__ __align__(16) uint32_t h[256 * 32];
    hist_range_lanes(h, in, n, chunk_bytes, spc, mult, seg_hist, blockIdx.x);
}

// ---------------------------------------------------------------------------------------------
int hzk_histogram(hz_ctx* ctx, const uint8_t* d_in, uint64_t n, uint32_t chunk_bytes, uint32_t K,
                  uint32_t* d_seg_hist) {
    if (K == 0) return HZ_OK;
    uint32_t spc = (chunk_bytes + HZ_SEG_BYTES - 1) / HZ_SEG_BYTES;
    uint64_t grid = (uint64_t)K * spc;
    if (grid > 0x7fffffffull) return hz_fail(ctx, HZ_ERR_ARG, "too many segments (%llu)", (unsigned long long)grid);
    const int variant = ctx->knobs.hist;
    if (variant == 0) {
        const size_t smem = (64 * 256 + 256) * sizeof(uint32_t);
        if (!ctx->attr_hist) {
            HZ_CUDA(ctx, cudaFuncSetAttribute(hist_seg_private, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            ctx->attr_hist = true;
        }
        HZ_LAUNCH(ctx, "hist_seg_private", hist_seg_private, (unsigned)grid, HZ_THREADS, smem,
                  d_in, n, chunk_bytes, spc, d_seg_hist);
    } else if (variant == 2) {
        uint32_t mult = hz_range_mult(spc, ctx->knobs.range_mult);
        if (ctx->knobs.hist_range == 0) mult = 1;     // developer knob: one segment per CTA
        const uint32_t rpc = (spc + mult - 1) / mult;
        HZ_LAUNCH(ctx, "hist_seg_lanes", hist_seg_lanes, (unsigned)((uint64_t)K * rpc), HZ_THREADS, 0,
                  d_in, n, chunk_bytes, spc, mult, d_seg_hist);
    } else {
        HZ_LAUNCH(ctx, "hist_seg_atomic", hist_seg_atomic, (unsigned)grid, HZ_THREADS, 0,
                  d_in, n, chunk_bytes, spc, d_seg_hist);
    }
    return HZ_OK;
}
