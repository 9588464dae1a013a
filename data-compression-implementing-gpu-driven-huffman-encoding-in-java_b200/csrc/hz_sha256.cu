// hz_sha256.cu — SHA-256 of every chunk on the device ("next").
//
// Replaces ChecksumUtil.computeSha256 per chunk (util/ChecksumUtil.java:11-27, called from
// cpu/CpuCompressionService.java:226-228 and :536).  SHA-256 is a Merkle–Damgård chain, so a
// chunk is inherently sequential: the only parallelism is ACROSS chunks.  One thread hashes one
// chunk; this pays off for many small chunks (the 64 KiB..4 MiB sweep) and is slower than the
// host's SHA units for a handful of 16-32 MiB chunks, which is why the file-level API hashes on
// the host while the GPU encodes (see hz_container.cpp).
#include "hz_common.cuh"

__constant__ uint32_t SHA_K[64] = {
    0x428a2f98,0x71374491,0xb5c0fbcf,0xe9b5dba5,0x3956c25b,0x59f111f1,0x923f82a4,0xab1c5ed5,
    0xd807aa98,0x12835b01,0x243185be,0x550c7dc3,0x72be5d74,0x80deb1fe,0x9bdc06a7,0xc19bf174,
    0xe49b69c1,0xefbe4786,0x0fc19dc6,0x240ca1cc,0x2de92c6f,0x4a7484aa,0x5cb0a9dc,0x76f988da,
    0x983e5152,0xa831c66d,0xb00327c8,0xbf597fc7,0xc6e00bf3,0xd5a79147,0x06ca6351,0x14292967,
    0x27b70a85,0x2e1b2138,0x4d2c6dfc,0x53380d13,0x650a7354,0x766a0abb,0x81c2c92e,0x92722c85,
    0xa2bfe8a1,0xa81a664b,0xc24b8b70,0xc76c51a3,0xd192e819,0xd6990624,0xf40e3585,0x106aa070,
    0x19a4c116,0x1e376c08,0x2748774c,0x34b0bcb5,0x391c0cb3,0x4ed8aa4a,0x5b9cca4f,0x682e6ff3,
    0x748f82ee,0x78a5636f,0x84c87814,0x8cc70208,0x90befffa,0xa4506ceb,0xbef9a3f7,0xc67178f2};

__device__ __forceinline__ uint32_t rotr32(uint32_t x, int n) { return __funnelshift_r(x, x, n); }

__device__ __forceinline__ void sha_block(uint32_t h[8], uint32_t w[16]) {
    uint32_t a=h[0],b=h[1],c=h[2],d=h[3],e=h[4],f=h[5],g=h[6],hh=h[7];
#pragma unroll
    for (int i = 0; i < 64; ++i) {
        if (i >= 16) {
            uint32_t w15 = w[(i + 1) & 15], w2 = w[(i + 14) & 15];
            uint32_t s0 = rotr32(w15, 7) ^ rotr32(w15, 18) ^ (w15 >> 3);
            uint32_t s1 = rotr32(w2, 17) ^ rotr32(w2, 19) ^ (w2 >> 10);
            w[i & 15] = w[i & 15] + s0 + w[(i + 9) & 15] + s1;
        }
        uint32_t S1 = rotr32(e, 6) ^ rotr32(e, 11) ^ rotr32(e, 25);
        uint32_t ch = (e & f) ^ (~e & g);
        uint32_t t1 = hh + S1 + ch + SHA_K[i] + w[i & 15];
        uint32_t S0 = rotr32(a, 2) ^ rotr32(a, 13) ^ rotr32(a, 22);
        uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
        uint32_t t2 = S0 + mj;
        hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
    }
    h[0]+=a; h[1]+=b; h[2]+=c; h[3]+=d; h[4]+=e; h[5]+=f; h[6]+=g; h[7]+=hh;
}

__global__ void __launch_bounds__(64)
sha256_chunks_kernel(const uint8_t* __restrict__ in, uint64_t n, uint32_t chunk_bytes, uint32_t K,
                     uint8_t* __restrict__ digests) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    const uint64_t beg = (uint64_t)k * chunk_bytes;
    const uint64_t len = n - beg < chunk_bytes ? n - beg : chunk_bytes;
    const uint8_t* p = in + beg;
    uint32_t h[8] = {0x6a09e667,0xbb67ae85,0x3c6ef372,0xa54ff53a,0x510e527f,0x9b05688c,0x1f83d9ab,0x5be0cd19};
    uint32_t w[16];
    const uint64_t full = len / 64;
    const bool aligned = (reinterpret_cast<uintptr_t>(p) & 15) == 0;
    for (uint64_t blk = 0; blk < full; ++blk) {
        const uint8_t* q = p + blk * 64;
        if (aligned) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                uint4 v = reinterpret_cast<const uint4*>(q)[i];
                w[4*i] = bswap32(v.x); w[4*i+1] = bswap32(v.y); w[4*i+2] = bswap32(v.z); w[4*i+3] = bswap32(v.w);
            }
        } else {
#pragma unroll
            for (int i = 0; i < 16; ++i)
                w[i] = (uint32_t)q[4*i] << 24 | (uint32_t)q[4*i+1] << 16 | (uint32_t)q[4*i+2] << 8 | q[4*i+3];
        }
        sha_block(h, w);
    }
    // padding
    const uint32_t rem = (uint32_t)(len - full * 64);
    uint8_t tail[128];
    for (uint32_t i = 0; i < rem; ++i) tail[i] = p[full * 64 + i];
    tail[rem] = 0x80;
    const uint32_t tl = (rem + 9 <= 64) ? 64 : 128;
    for (uint32_t i = rem + 1; i < tl; ++i) tail[i] = 0;
    const uint64_t bits = len * 8;
    for (int i = 0; i < 8; ++i) tail[tl - 1 - i] = (uint8_t)(bits >> (8 * i));
    for (uint32_t b = 0; b < tl; b += 64) {
        for (int i = 0; i < 16; ++i)
            w[i] = (uint32_t)tail[b+4*i] << 24 | (uint32_t)tail[b+4*i+1] << 16 | (uint32_t)tail[b+4*i+2] << 8 | tail[b+4*i+3];
        sha_block(h, w);
    }
    uint8_t* d = digests + (size_t)k * 32;
    for (int i = 0; i < 8; ++i) { d[4*i] = h[i] >> 24; d[4*i+1] = h[i] >> 16; d[4*i+2] = h[i] >> 8; d[4*i+3] = h[i]; }
}

int hzk_sha256(hz_ctx* ctx, const uint8_t* d_in, uint64_t n, uint32_t chunk_bytes, uint32_t K, uint8_t* d_digests) {
    if (K == 0) return HZ_OK;
    HZ_LAUNCH(ctx, "sha256_chunks", sha256_chunks_kernel, (K + 63) / 64, 64, 0, d_in, n, chunk_bytes, K, d_digests);
    return HZ_OK;
}
