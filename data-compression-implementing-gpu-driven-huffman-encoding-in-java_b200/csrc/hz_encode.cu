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
// no atomics on global memory, no pre-zeroed output.  Per tile of 8192 symbols:
//   1. one 256-bit coalesced load per thread (32 symbols), issued one tile ahead;
//   2. codeword gather from a bank-replicated shared-memory LUT (entry [sym][lane]: every lane
//      reads its own bank, conflict-free for any symbol distribution; the LUT sits on a 32 KiB
//      boundary of the shared window so that one LOP3 forms the address);
//   3. reduce-merge: pairs, then quads of codewords are folded into <=64-bit containers in
//      registers (arXiv 2010.10039 §IV-B);
//   4. exclusive scan of the per-thread bit counts (warp shuffles + one cross-warp step);
//   5. shuffle-merge replaced by its shared-memory equivalent: a thread walks its containers
//      once, assembling the words of its bit range in a register and OR-ing every word into the
//      staging buffer exactly once (red.shared.or; only the first and last word of a thread are
//      shared with its neighbours);
//   6. the completed 16-byte units are byte-swapped to the stream's MSB-first order and written
//      with aligned 128-bit stores; the trailing partial unit is carried to the next tile.
// The byte shared by two neighbouring segments is written by the LATER segment, which
// recomputes the previous segment's last <8 bits from its last 7 symbols.
// Chunks whose longest code exceeds 16 bits take the "wide" path (2 symbols per container,
// 8 symbols per thread per tile, unreplicated 64-bit LUT): correct for lengths up to 32, slower.
#include "hz_common.cuh"

#define ENC_SPT 32                                  // symbols per thread per tile (fast path)
#define ENC_TILE (HZ_THREADS * ENC_SPT)             // 8192
#define ENC_STAGE_WORDS (ENC_TILE * 16 / 32 + 16)   // 16 KiB of bits + carry unit + slack
#define ENC_LUT_BYTES (256 * 32 * 4)                // [sym][lane] uint32: len<<16 | code
#define ENC_LOW_BYTES (ENC_STAGE_WORDS * 4 + 128)    // staging buffer + scan scratch, below the LUT
#define ENC_SMEM_BYTES (65536 - 1024)               // the driver reserves the first 1 KiB of the shared window

static_assert(HZ_SEG_BYTES % ENC_TILE == 0, "a segment is a whole number of tiles");
static_assert(ENC_LOW_BYTES <= 32768 - 1024, "staging buffer must fit below the LUT");

__device__ __forceinline__ uint32_t enc_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t enc_lds32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ void enc_sts32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void enc_or32(uint32_t a, uint32_t v) { asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t enc_pin(uint32_t v) { asm volatile("" : "+r"(v)); return v; }
__device__ __forceinline__ uint32_t shl_c(uint32_t x, uint32_t s) { uint32_t r; asm("shl.b32 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(s)); return r; }   // s >= 32 -> 0

// 32 symbols (8 little-endian words) starting at q; nvalid of them exist
__device__ __forceinline__ void load_syms32(const uint8_t* q, int nvalid, uint32_t w[8]) {
    if (nvalid == 32 && (reinterpret_cast<uintptr_t>(q) & 31) == 0) {
        asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]) : "l"(q));
    } else if (nvalid == 32 && (reinterpret_cast<uintptr_t>(q) & 3) == 0) {
        const uint32_t* q4 = reinterpret_cast<const uint32_t*>(q);
#pragma unroll
        for (int i = 0; i < 8; ++i) w[i] = __ldg(q4 + i);
    } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) w[i] = 0;
        for (int i = 0; i < nvalid; ++i) w[i >> 2] |= (uint32_t)q[i] << (8 * (i & 3));
    }
}

// Block-wide exclusive scan of `v` with ONE barrier: every warp publishes its total, then each
// warp sums the totals of the warps before it.  warp_tot is double-buffered by the caller (`par`).
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t* warp_tot, uint32_t par, uint32_t* total) {
    const uint32_t t = threadIdx.x, lane = t & 31, wid = t >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t o = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += o;
    }
    uint32_t* wt = warp_tot + par * (HZ_THREADS / 32);
    if (lane == 31) wt[wid] = inc;
    __syncthreads();
    const uint32_t x = lane < HZ_THREADS / 32 ? wt[lane] : 0;
    const uint32_t before = __reduce_add_sync(0xffffffffu, lane < wid ? x : 0);
    *total = __reduce_add_sync(0xffffffffu, x);
    return before + inc - v;
}

// Running state of a thread that appends containers to the staging bit buffer.
struct Emit {
    uint32_t wp;     // shared address of the current (partially assembled) word
    uint32_t f;      // bits of that word that precede this thread's next bit
    uint32_t a0;     // this thread's bits of the current word
};
// Append the `len` (1..64) low bits of (ch:cl), MSB first.
__device__ __forceinline__ void emit_put(Emit& E, uint32_t ch, uint32_t cl, uint32_t len) {
    // left-align the container in 64 bits
    const uint32_t s = 64 - len;                               // 0..63
    uint32_t vh, vl;
    if (s >= 32) { vh = cl << (s - 32); vl = 0; }
    else { vh = __funnelshift_l(cl, ch, s); vl = cl << s; }
    // place it at bit f of the 96-bit window that starts at the current word
    const uint32_t w0 = vh >> E.f;
    const uint32_t w1 = __funnelshift_r(vl, vh, E.f);
    const uint32_t w2 = __funnelshift_r(0u, vl, E.f);
    E.a0 |= w0;
    const uint32_t e = E.f + len;
    if (e >= 32) {
        enc_or32(E.wp, E.a0); E.a0 = w1;
        if (e >= 64) { enc_or32(E.wp + 4, w1); E.a0 = w2; }
    }
    E.wp += (e >> 5) << 2;
    E.f = e & 31;
}

// ---------------------------------------------------------------------------------------------
// common tail of a tile: flush completed 16-byte units of the staging buffer, carry the rest
// ---------------------------------------------------------------------------------------------
struct SegOut {
    uint64_t unit0;      // absolute 16-byte unit index of stage word 0
    uint64_t own_lo;     // first byte address this CTA writes
    uint32_t cur;        // bits present in the staging buffer
};

__device__ __forceinline__ void flush_tile(uint32_t* stage, SegOut& O) {
    const uint32_t t = threadIdx.x;
    const uint32_t full = O.cur >> 7;
    for (uint32_t u = t; u < full; u += HZ_THREADS) {
        uint4 v = reinterpret_cast<uint4*>(stage)[u];
        reinterpret_cast<uint4*>(stage)[u] = make_uint4(0, 0, 0, 0);
        v.x = bswap32(v.x); v.y = bswap32(v.y); v.z = bswap32(v.z); v.w = bswap32(v.w);
        const uint64_t addr = (O.unit0 + u) * 16;
        if (addr >= O.own_lo) {
            *reinterpret_cast<uint4*>(addr) = v;
        } else {                                   // first unit of the segment: the bytes before
            const uint32_t wv[4] = {v.x, v.y, v.z, v.w};   // own_lo belong to the previous segment
            for (int b = 0; b < 16; ++b)
                if (addr + b >= O.own_lo)
                    *reinterpret_cast<uint8_t*>(addr + b) = (uint8_t)(wv[b >> 2] >> (8 * (b & 3)));
        }
    }
    if (full > 0) {                                // move the partial unit to the front
        __syncthreads();                           // unit 0 has been flushed and cleared
        if (t < 4) { const uint32_t carry = stage[full * 4 + t]; stage[full * 4 + t] = 0; stage[t] = carry; }
        O.unit0 += full;
        O.cur &= 127;
    }
    __syncthreads();
}

__device__ __forceinline__ void flush_tail(const uint32_t* stage, const SegOut& O, bool last_seg) {
    const uint32_t t = threadIdx.x;
    const uint32_t nb = last_seg ? (O.cur + 7) >> 3 : O.cur >> 3;
    if (t < nb) {
        const uint64_t addr = O.unit0 * 16 + t;
        if (addr >= O.own_lo)
            *reinterpret_cast<uint8_t*>(addr) = (uint8_t)(stage[t >> 2] >> (24 - 8 * (t & 3)));
    }
}

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(HZ_THREADS, 3)
encode_kernel(const uint8_t* __restrict__ in, uint64_t n, uint32_t chunk_bytes, uint32_t spc,
              const uint8_t* __restrict__ len_tab, const uint32_t* __restrict__ code_tab,
              const uint64_t* __restrict__ comp_off, const uint64_t* __restrict__ seg_bitoff,
              uint32_t K, uint8_t* __restrict__ out, uint64_t out_cap, int* status) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
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
    const uint8_t* p = in + cbeg + sbeg;
    // every start-up load is issued before the first use (they are independent)
    const uint32_t mylen = len_tab[(size_t)k * 256 + t];
    const uint32_t mycode = code_tab[(size_t)k * 256 + t];
    const uint64_t total_bytes = comp_off[K], chunk_off = comp_off[k], seg_off = seg_bitoff[seg];
    uint32_t w[8], wn[8];
    {
        const uint32_t first = t * ENC_SPT;
        const int nvalid = first >= slen ? 0 : (slen - first >= ENC_SPT ? ENC_SPT : (int)(slen - first));
        load_syms32(p + first, nvalid, wn);
    }
    if (total_bytes > out_cap) { if (t == 0) hz_set_status(status, HZ_ERR_OUT_TOO_SMALL); return; }

    // shared memory: [stage | scan scratch] ... [LUT on a 32 KiB boundary of the shared window]
    const uint32_t base_a = enc_smem_u32(smem_raw);
    const uint32_t lut_a = (base_a + ENC_LOW_BYTES + 32767u) & ~32767u;
    if (lut_a + ENC_LUT_BYTES > base_a + ENC_SMEM_BYTES) {      // cannot happen with the 1 KiB driver reservation
        if (t == 0) hz_set_status(status, HZ_ERR_CUDA);
        return;
    }
    uint32_t* stage = reinterpret_cast<uint32_t*>(smem_raw);
    uint32_t* warp_tot = stage + ENC_STAGE_WORDS;
    uint32_t* lut = reinterpret_cast<uint32_t*>(smem_raw + (lut_a - base_a));
    const uint32_t stage_a = enc_pin(base_a);
    uint32_t par = 0;

    // codebook of this chunk
    const bool wide = __syncthreads_or(mylen > 16);    // block-uniform: codes longer than 16 bits
    uint64_t* lut64 = reinterpret_cast<uint64_t*>(lut);
    if (!wide) {
        const uint32_t e = (mylen << 16) | mycode;
        uint4 e4 = make_uint4(e, e, e, e);
        uint4* row = reinterpret_cast<uint4*>(&lut[t * 32]);
#pragma unroll
        for (int i = 0; i < 8; ++i) row[i] = e4;
    } else {
        lut64[t] = ((uint64_t)mylen << 32) | mycode;
    }
    for (uint32_t i = t; i < ENC_STAGE_WORDS; i += HZ_THREADS) stage[i] = 0;
    __syncthreads();

    // absolute bit address of the segment's first bit, and the 16-byte unit it falls in
    const uint64_t P0 = chunk_off * 8 + seg_off;
    const uint64_t out_addr = reinterpret_cast<uint64_t>(out);
    const uint64_t G0 = out_addr * 8 + P0;
    SegOut O;
    O.unit0 = G0 >> 7;
    O.cur = (uint32_t)(G0 & 127);
    O.own_lo = out_addr + (P0 >> 3);

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
            Emit E; E.wp = stage_a + (((O.cur - r) >> 5) << 2); E.f = (O.cur - r) & 31; E.a0 = 0;
            emit_put(E, 0, bits, r);
            if (E.a0) enc_or32(E.wp, E.a0);
        }
    }

    if (!wide) {
        // ---- fast path: 32 symbols per thread per tile, software-pipelined loads ------------------
        const uint32_t lanebase = enc_pin(lut_a | (lane << 2));
        for (uint32_t tile = 0; tile < slen; tile += ENC_TILE) {
            const uint32_t first = tile + t * ENC_SPT;
            const int nvalid = first >= slen ? 0 : (slen - first >= ENC_SPT ? ENC_SPT : (int)(slen - first));
#pragma unroll
            for (int i = 0; i < 8; ++i) w[i] = wn[i];
            if (tile + ENC_TILE < slen) {                       // prefetch the next tile
                const uint32_t nf = first + ENC_TILE;
                const int nv = nf >= slen ? 0 : (slen - nf >= ENC_SPT ? ENC_SPT : (int)(slen - nf));
                load_syms32(p + nf, nv, wn);
            }
            // ---- gather + reduce-merge: 8 containers of 4 symbols -------------------------------
            uint32_t ch[8], cl[8], L[8];
            uint32_t tot = 0;
            const bool ragged = tile + ENC_TILE > slen;         // block-uniform
#pragma unroll
            for (int g = 0; g < 8; ++g) {
                const uint32_t x = w[g];
                uint32_t e0 = enc_lds32(((x << 7) & 0x7F80u) | lanebase);
                uint32_t e1 = enc_lds32(((x >> 1) & 0x7F80u) | lanebase);
                uint32_t e2 = enc_lds32(((x >> 9) & 0x7F80u) | lanebase);
                uint32_t e3 = enc_lds32(((x >> 17) & 0x7F80u) | lanebase);
                if (ragged) {                                   // last tile of a short segment
                    if (g * 4 + 0 >= nvalid) e0 = 0;
                    if (g * 4 + 1 >= nvalid) e1 = 0;
                    if (g * 4 + 2 >= nvalid) e2 = 0;
                    if (g * 4 + 3 >= nvalid) e3 = 0;
                }
                const uint32_t l1 = e1 >> 16, l3 = e3 >> 16;
                const uint32_t p01 = ((e0 & 0xFFFFu) << l1) | (e1 & 0xFFFFu);     // <= 32 bits
                const uint32_t p23 = ((e2 & 0xFFFFu) << l3) | (e3 & 0xFFFFu);
                const uint32_t L23 = (e2 >> 16) + l3, L01 = (e0 >> 16) + l1;      // <= 32 each
                ch[g] = __funnelshift_lc(p01, 0u, L23);                           // (p01 << L23) >> 32
                cl[g] = shl_c(p01, L23) | p23;
                L[g] = L01 + L23;
                tot += L[g];
            }
            // ---- exclusive scan of per-thread bit counts -----------------------------------------
            uint32_t tile_bits;
            const uint32_t off = O.cur + block_excl_scan(tot, warp_tot, par, &tile_bits);
            par ^= 1;
            // ---- every word of this thread's bit range is OR-ed into the staging buffer once -------
            Emit E; E.wp = stage_a + ((off >> 5) << 2); E.f = off & 31; E.a0 = 0;
            if (!ragged) {
#pragma unroll
                for (int g = 0; g < 8; ++g) emit_put(E, ch[g], cl[g], L[g]);
            } else {
#pragma unroll
                for (int g = 0; g < 8; ++g) if (L[g]) emit_put(E, ch[g], cl[g], L[g]);
            }
            if (E.a0) enc_or32(E.wp, E.a0);
            __syncthreads();
            O.cur += tile_bits;
            flush_tile(stage, O);
        }
    } else {
        // ---- wide path: codes of up to 32 bits, 8 symbols per thread per tile --------------------
        for (uint32_t tile = 0; tile < slen; tile += HZ_THREADS * 8) {
            const uint32_t first = tile + t * 8;
            const int nvalid = first >= slen ? 0 : (slen - first >= 8 ? 8 : (int)(slen - first));
            uint64_t c[4]; uint32_t L[4];
            uint32_t tot = 0;
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
            uint32_t tile_bits;
            const uint32_t off = O.cur + block_excl_scan(tot, warp_tot, par, &tile_bits);
            par ^= 1;
            Emit E; E.wp = stage_a + ((off >> 5) << 2); E.f = off & 31; E.a0 = 0;
#pragma unroll
            for (int g = 0; g < 4; ++g) if (L[g]) emit_put(E, (uint32_t)(c[g] >> 32), (uint32_t)c[g], L[g]);
            if (E.a0) enc_or32(E.wp, E.a0);
            __syncthreads();
            O.cur += tile_bits;
            flush_tile(stage, O);
        }
    }
    flush_tail(stage, O, last_seg);
}

int hzk_encode(hz_ctx* ctx, const uint8_t* d_in, uint64_t n, uint32_t chunk_bytes, uint32_t K,
               const uint8_t* d_len, const uint32_t* d_code, const uint64_t* d_comp_off,
               const uint64_t* d_seg_bitoff, uint8_t* d_out, uint64_t out_cap) {
    if (K == 0) return HZ_OK;
    uint32_t spc = (chunk_bytes + HZ_SEG_BYTES - 1) / HZ_SEG_BYTES;
    uint64_t grid = (uint64_t)K * spc;
    if (grid > 0x7fffffffull) return hz_fail(ctx, HZ_ERR_ARG, "too many segments");
    static bool attr_done = false;
    if (!attr_done) {
        HZ_CUDA(ctx, cudaFuncSetAttribute(encode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ENC_SMEM_BYTES));
        attr_done = true;
    }
    HZ_LAUNCH(ctx, "encode", encode_kernel, (unsigned)grid, HZ_THREADS, ENC_SMEM_BYTES,
              d_in, n, chunk_bytes, spc, d_len, d_code, d_comp_off, d_seg_bitoff, K, d_out, out_cap, ctx->d_status);
    return HZ_OK;
}
