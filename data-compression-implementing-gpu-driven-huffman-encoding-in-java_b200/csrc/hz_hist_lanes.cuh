// hz_hist_lanes.cuh — body of the lane-private histogram kernel (hist_seg_lanes, hz_hist.cu), shared with the chained
// histogram -> codebook kernel (hz_codebook.cu).
#pragma once
#include "hz_common.cuh"

// Histogram of range `bid` (ranges are numbered chunk by chunk: rpc = ceil(spc / mult) per chunk, `mult` segments of
// HZ_SEG_BYTES each = the unit an encoder group codes) into the range's first segment slot of seg_hist, zeros into
// its other slots.  One uint32 counter per (bin, lane): word bin*32 + lane, so a lane always hits its own bank and
// every shared-memory atomic of a warp is ONE wavefront whatever the byte distribution.  All HZ_THREADS threads;
// h = 32 KiB of shared memory.  Ends with the CTA's results stored (no barrier after the stores).
__device__ __forceinline__ void hist_range_lanes(uint32_t* h, const uint8_t* __restrict__ in, uint64_t n, uint32_t chunk_bytes,
                                                 uint32_t spc, uint32_t mult, uint32_t* __restrict__ seg_hist, uint32_t bid) {
    const uint32_t t = threadIdx.x, lane = t & 31;
    const uint32_t rpc = (spc + mult - 1) / mult;
    const uint32_t k = bid / rpc, s0 = (bid - k * rpc) * mult;
    const uint64_t cbeg = (uint64_t)k * chunk_bytes;
    const uint64_t clen = n - cbeg < chunk_bytes ? n - cbeg : chunk_bytes;
    const uint64_t rbeg = (uint64_t)s0 * HZ_SEG_BYTES;
    uint32_t* dst = seg_hist + ((size_t)k * spc + s0) * 256;
    const uint32_t nslots = spc - s0 < mult ? spc - s0 : mult;
    for (uint32_t j = 1; j < nslots; ++j) dst[j * 256 + t] = 0;
    if (rbeg >= clen) { dst[t] = 0; return; }
    const uint64_t sbeg = cbeg + rbeg;
    const uint64_t rl = clen - rbeg;
    const uint32_t slen = rl < (uint64_t)mult * HZ_SEG_BYTES ? (uint32_t)rl : mult * HZ_SEG_BYTES;
    {
        const uint4 z = make_uint4(0, 0, 0, 0);
        uint4* h4 = reinterpret_cast<uint4*>(h);
#pragma unroll
        for (int i = 0; i < 8; ++i) h4[t + i * HZ_THREADS] = z;
    }
    __syncthreads();
    const uint8_t* p = in + sbeg;
    uint32_t head = (uint32_t)((16 - (reinterpret_cast<uintptr_t>(p) & 15)) & 15);
    if (head > slen) head = slen;
    const uint32_t nvec = (slen - head) >> 4;
    const uint32_t tail = slen - head - (nvec << 4);
    const uint32_t mine = (uint32_t)__cvta_generic_to_shared(h) + lane * 4;      // this lane's column
    auto bump = [&](uint32_t sym) {
        asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(sym * 128u + mine) : "memory");
    };
    if (t < head) bump(p[t]);
    if (t < tail) bump(p[head + (nvec << 4) + t]);
    const uint4* pv = reinterpret_cast<const uint4*>(p + head);
#pragma unroll 2
    for (uint32_t i = t; i < nvec; i += HZ_THREADS) {
        const uint4 v = ld_stream_u4(pv + i);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            bump(__byte_perm(w[g], 0, 0x4440));
            bump(__byte_perm(w[g], 0, 0x4441));
            bump(__byte_perm(w[g], 0, 0x4442));
            bump(__byte_perm(w[g], 0, 0x4443));
        }
    }
    __syncthreads();
    uint32_t s = 0;
#pragma unroll 8
    for (uint32_t j = 0; j < 32; ++j) s += h[t * 32 + ((j + t) & 31)];
    dst[t] = s;
}
