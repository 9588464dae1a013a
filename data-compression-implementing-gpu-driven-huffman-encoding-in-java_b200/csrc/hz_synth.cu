// hz_synth.cu — deterministic synthetic byte streams for the bench / tests (NOT part of the
// drop-in boundary; declared in include/huffb200_synth.h).  byte i of the stream is
// qtable[ mix64(seed, i) >> 48 ], where qtable is a 65536-entry quantised inverse CDF supplied
// by the caller, so the same stream can be regenerated on the host with integer arithmetic.
#include "hz_common.cuh"
#include "../../include/huffb200_synth.h"

__host__ __device__ __forceinline__ uint64_t hz_mix64(uint64_t seed, uint64_t i) {
    uint64_t z = seed + (i + 1) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__global__ void __launch_bounds__(256)
synth_kernel(uint8_t* __restrict__ out, uint64_t n, uint64_t offset, uint64_t seed, const uint8_t* __restrict__ qtable) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x * 16;
    for (uint64_t base = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 16; base < n; base += stride) {
        if (base + 16 <= n && ((reinterpret_cast<uintptr_t>(out) + base) & 15) == 0) {
            uint32_t w[4];
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                uint32_t x = 0;
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    x |= (uint32_t)qtable[hz_mix64(seed, offset + base + g * 4 + j) >> 48] << (8 * j);
                w[g] = x;
            }
            *reinterpret_cast<uint4*>(out + base) = make_uint4(w[0], w[1], w[2], w[3]);
        } else {
            for (uint64_t i = base; i < n && i < base + 16; ++i) out[i] = qtable[hz_mix64(seed, offset + i) >> 48];
        }
    }
}

extern "C" int hz_synth_fill(hz_ctx* ctx, uint8_t* d_out, uint64_t n, uint64_t stream_offset, uint64_t seed,
                             const uint8_t* qtable65536) {
    if (!ctx || !qtable65536 || (n && !d_out)) return hz_fail(ctx, HZ_ERR_ARG, "hz_synth_fill: bad argument");
    HZ_CUDA(ctx, cudaSetDevice(ctx->device));
    if (n == 0) return HZ_OK;
    if (!hz_is_device_ptr(d_out)) return hz_fail(ctx, HZ_ERR_ARG, "hz_synth_fill: output must be device memory");
    const void* d_q = qtable65536;
    if (!hz_is_device_ptr(qtable65536)) {
        HZ_TRY(hz_reserve(ctx, &ctx->dec_misc, 65536));
        HZ_CUDA(ctx, cudaMemcpyAsync(ctx->dec_misc.p, qtable65536, 65536, cudaMemcpyHostToDevice, ctx->stream));
        d_q = ctx->dec_misc.p;
    }
    uint64_t blocks = (n / 16 + 255) / 256;
    if (blocks > (uint64_t)ctx->sm_count * 32) blocks = (uint64_t)ctx->sm_count * 32;
    if (blocks == 0) blocks = 1;
    HZ_LAUNCH(ctx, "synth_fill", synth_kernel, (unsigned)blocks, 256, 0, d_out, n, stream_offset, seed, (const uint8_t*)d_q);
    return HZ_OK;
}

extern "C" int hz_dev_reload_knobs(hz_ctx* ctx) {
    if (!ctx) return HZ_ERR_ARG;
    ctx->knobs = hz_knobs();
    hz_read_knobs(&ctx->knobs);
    return HZ_OK;
}
