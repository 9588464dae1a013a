// hz_encode.cu — bit-packing encoder (stage 3 of the encode pipeline).
//
// Replaces CpuCompressionService.encodeChunk + BitOutputStream.writeBits
// (service/cpu/CpuCompressionService.java:303-315, :711-737) and the TornadoVM packet kernel
// (service/gpu/TornadoKernels.java:115-205): concatenate code[sym] MSB-first for every input
// byte of a chunk, zero-pad the chunk's last byte, chunks back to back.
//
// One CTA encodes one segment (HZ_SEG_BYTES symbols of one chunk).  The absolute output bit
// offset of the segment is known up front (comp_off[chunk]*8 + seg_bitoff[segment], both derived
// from the histograms by the codebook kernel), so segments are independent: no look-back chain,
// no atomics on global memory, no pre-zeroed output.  Per tile of 4096 symbols:
//   1. 128-bit coalesced symbol loads (16 symbols per thread);
//   2. codeword gather from a bank-replicated shared-memory LUT (entry [sym][lane]: every lane
//      reads its own bank, conflict-free for any symbol distribution);
//   3. reduce-merge: each thread folds 4 consecutive codewords into one <=64-bit container
//      (arXiv 2010.10039 §IV-B, done in registers);
//   4. exclusive scan of the per-thread bit counts (warp shuffles + one cross-warp step);
//   5. shuffle-merge replaced by its shared-memory equivalent: every container is OR-ed at its
//      bit offset into a staging buffer that mirrors the output's 16-byte alignment;
//   6. the completed 16-byte units are byte-swapped to the stream's MSB-first order and written
//      with aligned 128-bit stores; the trailing partial unit is carried to the next tile.
// The byte shared by two neighbouring segments is written by the LATER segment, which
// recomputes the previous segment's last <8 bits from its last 7 symbols.
// Chunks whose longest code exceeds 16 bits take the "wide" path (2 symbols per container,
// 8 symbols per thread per tile, unreplicated 64-bit LUT): correct for lengths up to 32, slower.
#include "hz_common.cuh"

#define ENC_SPT 16                                  // symbols per thread per tile (fast path)
#define ENC_TILE (HZ_THREADS * ENC_SPT)             // 4096
#define ENC_STAGE_WORDS (ENC_TILE * 16 / 32 + 16)   // 8 KiB of bits + carry unit + slack
#define ENC_LUT_WORDS (256 * 32)

struct EncSmem {
    uint32_t lut[ENC_LUT_WORDS];        // fast path: [sym][lane] = len<<16 | code (len <= 16)
    uint32_t stage[2][ENC_STAGE_WORDS];
    uint32_t warp_tot[HZ_THREADS / 32 + 1];
};

// OR a right-aligned `len`-bit value into the staging bit buffer at bit position `bitpos`
// (bit 0 of the buffer = MSB of word 0).
__device__ __forceinline__ void stage_put(uint32_t* stage, uint32_t bitpos, uint64_t val, uint32_t len) {
    if (len == 0) return;
    uint64_t v = val << (64 - len);
    uint32_t hi = (uint32_t)(v >> 32), lo = (uint32_t)v;
    uint32_t wi = bitpos >> 5, sh = bitpos & 31;
    uint32_t w0 = hi >> sh;
    uint32_t w1 = __funnelshift_r(lo, hi, sh);
    uint32_t w2 = __funnelshift_r(0u, lo, sh);
    if (w0) atomicOr(&stage[wi], w0);
    if (w1) atomicOr(&stage[wi + 1], w1);
    if (w2) atomicOr(&stage[wi + 2], w2);
}

// Load up to 16 symbols starting at q (nvalid of them exist) into 4 little-endian words.
__device__ __forceinline__ void load_syms16(const uint8_t* q, int nvalid, uint32_t w[4]) {
    if (nvalid == 16 && (reinterpret_cast<uintptr_t>(q) & 15) == 0) {
        uint4 v = ld_stream_u4(reinterpret_cast<const uint4*>(q));
        w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
    } else {
        w[0] = w[1] = w[2] = w[3] = 0;
        for (int i = 0; i < nvalid; ++i) w[i >> 2] |= (uint32_t)q[i] << (8 * (i & 3));
    }
}

// Block-wide exclusive scan of `v`; returns this thread's exclusive prefix, *total = block sum.
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t* warp_tot, uint32_t* total) {
    const uint32_t t = threadIdx.x, lane = t & 31, wid = t >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t o = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += o;
    }
    if (lane == 31) warp_tot[wid] = inc;
    __syncthreads();
    if (t < 32) {
        uint32_t x = t < HZ_THREADS / 32 ? warp_tot[t] : 0;
        uint32_t xi = x;
#pragma unroll
        for (int d = 1; d < HZ_THREADS / 32; d <<= 1) {
            uint32_t o = __shfl_up_sync(0xffffffffu, xi, d);
            if (lane >= d) xi += o;
        }
        if (t < HZ_THREADS / 32) warp_tot[t] = xi - x;
        if (t == HZ_THREADS / 32 - 1) warp_tot[HZ_THREADS / 32] = xi;
    }
    __syncthreads();
    *total = warp_tot[HZ_THREADS / 32];
    return warp_tot[wid] + inc - v;
}

__global__ void __launch_bounds__(HZ_THREADS, 4)
encode_kernel(const uint8_t* __restrict__ in, uint64_t n, uint32_t chunk_bytes, uint32_t spc,
              const uint8_t* __restrict__ len_tab, const uint32_t* __restrict__ code_tab,
              const uint64_t* __restrict__ comp_off, const uint64_t* __restrict__ seg_bitoff,
              uint32_t K, uint8_t* __restrict__ out, uint64_t out_cap, int* status) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    EncSmem& S = *reinterpret_cast<EncSmem*>(smem_raw);
    const uint32_t t = threadIdx.x, lane = t & 31;
    const uint32_t seg = blockIdx.x;
    const uint32_t k = seg / spc, s = seg - k * spc;

    // geometry
    const uint64_t cbeg = (uint64_t)k * chunk_bytes;
    const uint64_t clen = n - cbeg < chunk_bytes ? n - cbeg : chunk_bytes;
    const uint64_t sbeg = (uint64_t)s * HZ_SEG_BYTES;
    if (sbeg >= clen) return;
    const uint32_t slen = (uint32_t)(clen - sbeg < HZ_SEG_BYTES ? clen - sbeg : HZ_SEG_BYTES);
    const bool last_seg = sbeg + slen >= clen;
    if (comp_off[K] > out_cap) { if (t == 0) hz_set_status(status, HZ_ERR_OUT_TOO_SMALL); return; }
    const uint8_t* p = in + cbeg + sbeg;

    // codebook of this chunk
    const uint32_t mylen = len_tab[(size_t)k * 256 + t];
    const uint32_t mycode = code_tab[(size_t)k * 256 + t];
    const bool wide = __syncthreads_or(mylen > 16);    // block-uniform: codes longer than 16 bits
    uint64_t* lut64 = reinterpret_cast<uint64_t*>(S.lut);
    if (!wide) {
        const uint32_t e = (mylen << 16) | mycode;
        uint4 e4 = make_uint4(e, e, e, e);
        uint4* row = reinterpret_cast<uint4*>(&S.lut[t * 32]);
#pragma unroll
        for (int i = 0; i < 8; ++i) row[i] = e4;
    } else {
        lut64[t] = ((uint64_t)mylen << 32) | mycode;
    }
    for (uint32_t i = t; i < 2 * ENC_STAGE_WORDS; i += HZ_THREADS) (&S.stage[0][0])[i] = 0;
    __syncthreads();

    // absolute bit address of the segment's first bit, and the 16-byte unit it falls in
    const uint64_t P0 = comp_off[k] * 8 + seg_bitoff[seg];
    const uint64_t out_addr = reinterpret_cast<uint64_t>(out);
    const uint64_t G0 = out_addr * 8 + P0;
    uint64_t unit0 = G0 >> 7;                          // absolute 16-byte unit index of stage word 0
    uint32_t cur = (uint32_t)(G0 & 127);               // bits already present in the staging buffer
    const uint64_t own_lo = out_addr + (P0 >> 3);      // first byte address this CTA writes
    int buf = 0;

    // the leading shared byte: previous segment's last (P0 & 7) bits, from its last 7 symbols
    if (t == 0) {
        const uint32_t r = (uint32_t)(P0 & 7);
        if (r) {
            uint32_t bits = 0, have = 0;
            for (int j = 1; j <= 7 && have < r; ++j) {
                uint32_t sym = p[-j];
                uint32_t l = len_tab[(size_t)k * 256 + sym];
                uint32_t c = code_tab[(size_t)k * 256 + sym];
                uint32_t take = l < 8 ? l : 8;
                bits |= (c & ((1u << take) - 1)) << have;
                have += take;
            }
            bits &= (1u << r) - 1;
            stage_put(S.stage[0], cur - r, bits, r);
        }
    }

    const uint32_t tile_syms = wide ? ENC_TILE / 2 : ENC_TILE;
    for (uint32_t tile = 0; tile < slen; tile += tile_syms) {
        uint32_t* stage = S.stage[buf];
        // ---- 1-3. load, gather, reduce-merge ----------------------------------------------------
        uint64_t c[4]; uint32_t L[4];
        uint32_t tot = 0;
        if (!wide) {
            const uint32_t first = tile + t * ENC_SPT;
            const int nvalid = first >= slen ? 0 : (slen - first >= ENC_SPT ? ENC_SPT : (int)(slen - first));
            uint32_t w[4];
            load_syms16(p + first, nvalid, w);
            const uint32_t* lutl = S.lut + lane;
            if (nvalid == ENC_SPT) {
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    uint64_t cc = 0; uint32_t ll = 0;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        uint32_t e = lutl[((w[g] >> (8 * j)) & 0xFF) * 32];
                        uint32_t l = e >> 16;
                        cc = (cc << l) | (e & 0xFFFF);
                        ll += l;
                    }
                    c[g] = cc; L[g] = ll; tot += ll;
                }
            } else {
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    uint64_t cc = 0; uint32_t ll = 0;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        if (g * 4 + j < nvalid) {
                            uint32_t e = lutl[((w[g] >> (8 * j)) & 0xFF) * 32];
                            uint32_t l = e >> 16;
                            cc = (cc << l) | (e & 0xFFFF);
                            ll += l;
                        }
                    }
                    c[g] = cc; L[g] = ll; tot += ll;
                }
            }
        } else {
            const uint32_t first = tile + t * 8;
            const int nvalid = first >= slen ? 0 : (slen - first >= 8 ? 8 : (int)(slen - first));
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                uint64_t cc = 0; uint32_t ll = 0;
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    if (g * 2 + j < nvalid) {
                        uint64_t e = lut64[p[first + g * 2 + j]];
                        uint32_t l = (uint32_t)(e >> 32);
                        cc = (cc << l) | (uint32_t)e;
                        ll += l;
                    }
                }
                c[g] = cc; L[g] = ll; tot += ll;
            }
        }
        // ---- 4. exclusive scan of per-thread bit counts -----------------------------------------
        uint32_t tile_bits;
        uint32_t off = cur + block_excl_scan(tot, S.warp_tot, &tile_bits);
        // ---- 5. merge into the staging buffer ---------------------------------------------------
#pragma unroll
        for (int g = 0; g < 4; ++g) { stage_put(stage, off, c[g], L[g]); off += L[g]; }
        __syncthreads();
        cur += tile_bits;
        // ---- 6. flush completed 16-byte units, carry the partial one -----------------------------
        const uint32_t full = cur >> 7;
        for (uint32_t u = t; u < full; u += HZ_THREADS) {
            uint4 v = reinterpret_cast<uint4*>(stage)[u];
            reinterpret_cast<uint4*>(stage)[u] = make_uint4(0, 0, 0, 0);
            v.x = bswap32(v.x); v.y = bswap32(v.y); v.z = bswap32(v.z); v.w = bswap32(v.w);
            const uint64_t addr = (unit0 + u) * 16;
            if (addr >= own_lo) {
                *reinterpret_cast<uint4*>(addr) = v;
            } else {                                   // first unit of the segment: the bytes before
                const uint32_t wv[4] = {v.x, v.y, v.z, v.w};   // own_lo belong to the previous segment
                for (int b = 0; b < 16; ++b)
                    if (addr + b >= own_lo)
                        *reinterpret_cast<uint8_t*>(addr + b) = (uint8_t)(wv[b >> 2] >> (8 * (b & 3)));
            }
        }
        if (full > 0) {
            if (t < 4) {                               // move the partial unit to the other buffer
                S.stage[buf ^ 1][t] = stage[full * 4 + t];
                stage[full * 4 + t] = 0;
            }
            buf ^= 1;
            unit0 += full;
            cur &= 127;
        }
        __syncthreads();
    }

    // ---- tail: the bytes of the last partial unit -----------------------------------------------
    {
        const uint32_t* stage = S.stage[buf];
        const uint32_t nb = last_seg ? (cur + 7) >> 3 : cur >> 3;
        if (t < nb) {
            const uint64_t addr = unit0 * 16 + t;
            if (addr >= own_lo)
                *reinterpret_cast<uint8_t*>(addr) = (uint8_t)(stage[t >> 2] >> (24 - 8 * (t & 3)));
        }
    }
}

int hzk_encode(hz_ctx* ctx, const uint8_t* d_in, uint64_t n, uint32_t chunk_bytes, uint32_t K,
               const uint8_t* d_len, const uint32_t* d_code, const uint64_t* d_comp_off,
               const uint64_t* d_seg_bitoff, uint8_t* d_out, uint64_t out_cap) {
    if (K == 0) return HZ_OK;
    uint32_t spc = (chunk_bytes + HZ_SEG_BYTES - 1) / HZ_SEG_BYTES;
    uint64_t grid = (uint64_t)K * spc;
    if (grid > 0x7fffffffull) return hz_fail(ctx, HZ_ERR_ARG, "too many segments");
    const size_t smem = sizeof(EncSmem);
    HZ_CUDA(ctx, cudaFuncSetAttribute(encode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    HZ_LAUNCH(ctx, "encode", encode_kernel, (unsigned)grid, HZ_THREADS, smem,
              d_in, n, chunk_bytes, spc, d_len, d_code, d_comp_off, d_seg_bitoff, K, d_out, out_cap, ctx->d_status);
    return HZ_OK;
}
