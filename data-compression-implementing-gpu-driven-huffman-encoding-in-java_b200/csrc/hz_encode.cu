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
//   A. one 256-bit coalesced load per thread (32 symbols), issued one tile ahead; every symbol is
//      looked up in a bank-replicated shared-memory LUT (entry [sym][lane] = codeword left-aligned
//      in bits 31.. | length in bits 4..0: every lane reads its own bank, conflict-free for any
//      symbol distribution) and appended to a 64-bit register accumulator with two funnel shifts
//      that take both the code bits and the shift amount straight from the LUT entry
//      (reduce-merge of arXiv 2010.10039 §IV-B, done sequentially in registers).  After every
//      second symbol the accumulator's completed 32-bit word, if any, goes to the thread's
//      PRIVATE staging row (17 words, odd stride): the thread's 32 codewords become one
//      contiguous, word-aligned bit string without knowing where it will land;
//   S. exclusive scan of the per-thread bit counts (warp shuffles + one cross-warp step, ONE barrier);
//   B. shuffle-merge: the thread funnel-shifts its private words to its bit offset in the dense
//      tile buffer.  Interior words are plain stores; the word a thread shares with its
//      predecessor is completed by a warp shuffle of the predecessor's tail (a thread always owns
//      >= 32 bits, so at most two threads meet in a word); only the two words at a warp's ends are
//      OR-ed with red.shared;
//   F. the completed 16-byte units are byte-swapped to the stream's MSB-first order and written
//      with aligned 128-bit stores; the trailing partial unit is carried to the next tile.
// Two barriers per tile.  A CTA is TWO independent 256-thread groups (named barriers), each encoding
// its own segment of the same chunk; they share only the 32 KiB LUT, which buys 32 resident warps
// per SM instead of 24.  The byte shared by two neighbouring segments is written by the LATER
// segment, which recomputes the previous segment's last <8 bits from its last 7 symbols.
// Chunks whose longest code has 17..27 bits take the "medium" instantiation (16 symbols per thread
// per tile, the same LUT entry format, a completed-word check after EVERY symbol, OR-ing merge);
// chunks whose longest code exceeds 27 bits take the "wide" instantiation (8 symbols per thread
// per tile, 64-bit LUT entries, a completed-word check after every symbol, every dense word
// OR-ed): correct for lengths up to 32, slower; it also serves the ragged last tile of a segment.
#include <cstdlib>
#include "hz_common.cuh"

#define ENC_SPT 32                                  // symbols per thread per tile (fast path)
#define ENC_TILE (HZ_THREADS * ENC_SPT)             // 8192
#define ENC_MED_SPT 16                              // symbols per thread per tile (codes of 17..27 bits)
#define ENC_MED_TILE (HZ_THREADS * ENC_MED_SPT)     // 4096
#define ENC_WIDE_SPT 8
#define ENC_WIDE_TILE (HZ_THREADS * ENC_WIDE_SPT)   // 2048
#define ENC_GROUPS 2                                // independent 256-thread groups per CTA
#define ENC_CTA (ENC_GROUPS * HZ_THREADS)
#define ENC_PRIV_STRIDE 17                          // words per thread: 16 complete + 1 partial / zero pad
#define ENC_DENSE_WORDS (ENC_TILE * 16 / 32 + 16)   // 16 KiB of bits + carry unit + slack
#define ENC_LUT_BYTES (256 * 32 * 4)                // [sym][lane] uint32
// per-group region
#define ENC_G_DENSE 0
#define ENC_G_WTOT (ENC_DENSE_WORDS * 4)
#define ENC_G_PRIV (ENC_G_WTOT + 64)
#define ENC_G_BYTES (ENC_G_PRIV + HZ_THREADS * ENC_PRIV_STRIDE * 4)
#define ENC_SMEM_BYTES (ENC_LUT_BYTES + ENC_GROUPS * ENC_G_BYTES)

static_assert(HZ_SEG_BYTES % ENC_TILE == 0, "a segment is a whole number of tiles");
static_assert(ENC_WIDE_TILE * 32 / 32 + 16 <= ENC_DENSE_WORDS, "wide tile must fit the dense buffer");
static_assert(ENC_MED_TILE * 27 / 32 + 16 <= ENC_DENSE_WORDS, "medium tile must fit the dense buffer");
static_assert(ENC_MED_SPT * 27 <= (ENC_PRIV_STRIDE - 2) * 32, "medium row must fit the private row");
static_assert(ENC_G_PRIV % 16 == 0 && ENC_G_BYTES % 16 == 0 && ENC_DENSE_WORDS % 4 == 0, "alignment");
static_assert(2 * (ENC_SMEM_BYTES + 1024) <= 227 * 1024, "two CTAs per SM");

__device__ __forceinline__ uint32_t enc_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t enc_lds32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ void enc_sts32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void enc_or32(uint32_t a, uint32_t v) { asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t enc_pin(uint32_t v) { asm volatile("" : "+r"(v)); return v; }
__device__ __forceinline__ uint32_t shl_c(uint32_t x, uint32_t s) { uint32_t r; asm("shl.b32 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(s)); return r; }   // s >= 32 -> 0
// barrier of one 256-thread group (ids 1..ENC_GROUPS; 0 is __syncthreads)
__device__ __forceinline__ void group_sync(uint32_t id) {
    if (id == 1) asm volatile("bar.sync 1, %0;" ::"n"(HZ_THREADS) : "memory");
    else asm volatile("bar.sync 2, %0;" ::"n"(HZ_THREADS) : "memory");
}
static_assert(ENC_GROUPS == 2, "group_sync names two barriers");

// 32 symbols (8 little-endian words) starting at q.  mode: 0 = 32-byte aligned, 1 = 4-byte aligned,
// 2 = bytes (nvalid of them exist).  Static register indices only: w[] must stay in registers.
__device__ __forceinline__ void load_syms32(const uint8_t* q, int mode, int nvalid, uint32_t w[8]) {
    if (mode == 0) {
        asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]) : "l"(q));
    } else if (mode == 1) {
        const uint32_t* q4 = reinterpret_cast<const uint32_t*>(q);
#pragma unroll
        for (int i = 0; i < 8; ++i) w[i] = __ldg(q4 + i);
    } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            uint32_t v = 0;
#pragma unroll
            for (int b = 0; b < 4; ++b)
                if (i * 4 + b < nvalid) v |= (uint32_t)q[i * 4 + b] << (8 * b);
            w[i] = v;
        }
    }
}

// Group-wide exclusive scan of `v` with ONE barrier: every warp publishes its total, then each
// warp sums the totals of the warps before it.  warp_tot is double-buffered by the caller (`par`).
__device__ __forceinline__ uint32_t group_excl_scan(uint32_t v, uint32_t* warp_tot, uint32_t par, uint32_t bar_id,
                                                    uint32_t lane, uint32_t wid, uint32_t* total) {
    uint32_t inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t o = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += o;
    }
    uint32_t* wt = warp_tot + par * (HZ_THREADS / 32);
    if (lane == 31) wt[wid] = inc;
    group_sync(bar_id);
    const uint32_t x = lane < HZ_THREADS / 32 ? wt[lane] : 0;
    const uint32_t before = __reduce_add_sync(0xffffffffu, lane < wid ? x : 0);
    *total = __reduce_add_sync(0xffffffffu, x);
    return before + inc - v;
}

// ---------------------------------------------------------------------------------------------
// state of the segment's output: the dense tile buffer holds bits [0, cur); dense unit 0 is the
// 16-byte unit *gptr of the output
// ---------------------------------------------------------------------------------------------
struct SegOut {
    uint4* gptr;         // output address of dense unit 0 (16-byte aligned)
    uint32_t skip;       // leading bytes of that unit owned by the previous segment (first flush only)
    uint32_t cur;        // bits present in the dense buffer
};

// phase F: write the completed 16-byte units, clear them, move the partial unit to the front.
// No barrier inside: the caller's next barrier (the scan of the next tile) orders it against the
// next phase B.  Thread 0 of the group alone touches unit 0 and the partial unit.
__device__ __forceinline__ void flush_tile(uint32_t* dense, SegOut& O, uint32_t tg) {
    const uint32_t full = O.cur >> 7;
    uint4* d4 = reinterpret_cast<uint4*>(dense);
    for (uint32_t u = tg; u < full; u += HZ_THREADS) {
        uint4 v = d4[u];
        d4[u] = make_uint4(0, 0, 0, 0);
        v.x = bswap32(v.x); v.y = bswap32(v.y); v.z = bswap32(v.z); v.w = bswap32(v.w);
        if (u != 0 || O.skip == 0) {
            O.gptr[u] = v;
        } else {                                   // first unit of the segment: its first `skip` bytes
            uint8_t* g = reinterpret_cast<uint8_t*>(O.gptr);   // belong to the previous segment
#pragma unroll
            for (int b = 0; b < 16; ++b) {
                const uint32_t wv = b < 4 ? v.x : (b < 8 ? v.y : (b < 12 ? v.z : v.w));
                if (b >= (int)O.skip) g[b] = (uint8_t)(wv >> (8 * (b & 3)));
            }
        }
    }
    if (full > 0) {
        if (tg == 0) {                             // after its own flush of unit 0 (program order)
            const uint4 c = d4[full];
            d4[full] = make_uint4(0, 0, 0, 0);
            d4[0] = c;
        }
        O.gptr += full;
        O.cur &= 127;
        O.skip = 0;
    }
}

__device__ __forceinline__ void flush_tail(const uint32_t* dense, const SegOut& O, bool last_seg, uint32_t tg) {
    const uint32_t nb = last_seg ? (O.cur + 7) >> 3 : O.cur >> 3;
    if (tg < nb && tg >= O.skip)
        reinterpret_cast<uint8_t*>(O.gptr)[tg] = (uint8_t)(dense[tg >> 2] >> (24 - 8 * (tg & 3)));
}

// phase B, general flavour: OR every non-zero word of the thread's bit string into the dense buffer.
// priv_a: shared address of the thread's private row (zero word after the last one), n bits.
__device__ __forceinline__ void merge_or(uint32_t priv_a, uint32_t dense_a, uint32_t off, uint32_t n) {
    if (n == 0) return;
    const uint32_t s = off & 31;
    const uint32_t d = dense_a + ((off >> 3) & ~3u);
    const uint32_t nd = (s + n + 31) >> 5;
    uint32_t prev = 0;
    for (uint32_t j = 0; j < nd; ++j) {
        const uint32_t w = enc_lds32(priv_a + 4 * j);
        const uint32_t o = __funnelshift_r(w, prev, s);
        prev = w;
        if (o) enc_or32(d + 4 * j, o);
    }
}

// phase B, fast flavour: requires n >= 32 for every thread of the tile (true whenever all 32
// symbols exist: every present symbol has a code of >= 1 bit), so a dense word holds bits of at
// most two threads.  The loop bound is warp-uniform, loads and stores use immediate offsets.
__device__ __forceinline__ void merge_shuffle(uint32_t priv_a, uint32_t dense_a, uint32_t off, uint32_t n, uint32_t lane) {
    const uint32_t s = off & 31;
    const uint32_t d = dense_a + ((off >> 3) & ~3u);
    const uint32_t e = s + n;
    const uint32_t nd = (e + 31) >> 5;                   // dense words touched
    const bool partial = (e & 31) != 0;                  // the last touched word is shared with the successor
    const uint32_t nst = nd - (partial ? 1u : 0u);       // words 0 .. nst-1 are stored by this thread
    const uint32_t jmax = __reduce_max_sync(0xffffffffu, nst);
    uint32_t prev = enc_lds32(priv_a);
    const uint32_t d0 = prev >> s;                       // first word: completed by the predecessor's tail
#pragma unroll
    for (int j = 1; j < ENC_PRIV_STRIDE; ++j) {
        if ((j & 1) && j >= (int)jmax) break;            // warp-uniform, tested every other word
        const uint32_t w = enc_lds32(priv_a + 4 * j);    // (lanes past their own row end read stale words: unused)
        const uint32_t o = __funnelshift_r(w, prev, s);
        prev = w;
        if (j < (int)nst) enc_sts32(d + 4 * j, o);
    }
    // the word this thread shares with its successor (index nd-1), branch-free; with n >= 32 a
    // partial last word implies nd >= 2 (otherwise the loads hit a neighbouring row: unused)
    const uint32_t ta = priv_a + 4 * nd;
    const uint32_t tw = __funnelshift_r(enc_lds32(ta - 4), enc_lds32(ta - 8), s);
    const uint32_t tail = partial ? tw : 0u;
    const uint32_t pt = __shfl_up_sync(0xffffffffu, tail, 1);
    // first word: plain store completed with the predecessor's tail; the two words at a warp's
    // ends are OR-ed (they are shared with another warp)
    asm volatile("{\n"
                 ".reg .pred p0, p31;\n"
                 "setp.eq.u32 p0, %4, 0;\n"
                 "setp.eq.u32 p31, %4, 31;\n"
                 "@p0 red.shared.or.b32 [%0], %1;\n"
                 "@!p0 st.shared.u32 [%0], %2;\n"
                 "@p31 red.shared.or.b32 [%3], %5;\n"
                 "}\n" ::"r"(d), "r"(d0), "r"(d0 | pt), "r"(d + 4 * (nd - 1)), "r"(lane), "r"(tail) : "memory");
}

// ---------------------------------------------------------------------------------------------
// phase A, fast path: append the two codewords of LUT entries e0, e1 to (hi:lo); when a 32-bit
// word completes, store it to the private row.  nb counts bits (low 16 bits exact; the upper
// bits collect code bits and are never read).
// ---------------------------------------------------------------------------------------------
struct Acc { uint32_t hi, lo, nb, ptr; };

__device__ __forceinline__ void acc_pair(Acc& a, uint32_t e0, uint32_t e1) {
    a.hi = __funnelshift_l(a.lo, a.hi, e0); a.lo = __funnelshift_l(e0, a.lo, e0);
    a.hi = __funnelshift_l(a.lo, a.hi, e1); a.lo = __funnelshift_l(e1, a.lo, e1);
    const uint32_t nb2 = a.nb + e0 + e1;
    if ((a.nb ^ nb2) & 32) {                                 // <= 32 bits were added: at most one word completes
        enc_sts32(a.ptr, __funnelshift_r(a.lo, a.hi, nb2));  // the 32 bits above the (nb2 & 31) pending ones
        a.ptr += 4;
    }
    a.nb = nb2;
}

// LUT address of byte J of x: lut[sym][lane]
template <int J>
__device__ __forceinline__ uint32_t lut_addr(uint32_t x, uint32_t lanebase) {
    // (taking the byte on the FMA pipe instead - IMAD.HI by a power of two - was measured on B200 and rejected:
    //  encode_kernel 2.174 -> 2.163 ms for bytes 0 and 3, 2.275 ms for all four)
    const uint32_t sym = __byte_perm(x, 0, 0x4440 + J);
    uint32_t a;
    asm("mad.lo.u32 %0, %1, 128, %2;" : "=r"(a) : "r"(sym), "r"(lanebase));
    return a;
}

template <bool RAGGED>
__device__ __forceinline__ uint32_t encode_thread32(const uint32_t w[8], int nvalid, uint32_t lanebase, uint32_t priv_a) {
    Acc a; a.hi = 0; a.lo = 0; a.nb = 0; a.ptr = priv_a;
#pragma unroll
    for (int g = 0; g < 8; ++g) {
        const uint32_t x = w[g];
        uint32_t e0 = enc_lds32(lut_addr<0>(x, lanebase));
        uint32_t e1 = enc_lds32(lut_addr<1>(x, lanebase));
        uint32_t e2 = enc_lds32(lut_addr<2>(x, lanebase));
        uint32_t e3 = enc_lds32(lut_addr<3>(x, lanebase));
        if (RAGGED) {
            if (g * 4 + 0 >= nvalid) e0 = 0;
            if (g * 4 + 1 >= nvalid) e1 = 0;
            if (g * 4 + 2 >= nvalid) e2 = 0;
            if (g * 4 + 3 >= nvalid) e3 = 0;
        }
        acc_pair(a, e0, e1);
        acc_pair(a, e2, e3);
    }
    const uint32_t nbits = a.nb & 0xFFFFu;
    const uint32_t v = nbits & 31;
    enc_sts32(a.ptr, shl_c(a.lo, 32 - v));                   // pending bits, left-aligned (0 when v == 0)
    if (v) enc_sts32(a.ptr + 4, 0);                          // zero word after the last one
    return nbits;
}

// phase A, medium path: up to 16 symbols (4 words), codes of up to 27 bits.  Same LUT entries
// (code left-aligned | length in bits 4..0), but the code may reach down to bit 5, so the length is
// masked before it is counted, and a 32-bit word can complete after every symbol.
__device__ __forceinline__ uint32_t encode_thread_med(const uint32_t w[4], int nvalid, uint32_t lanebase, uint32_t priv_a) {
    uint32_t hi = 0, lo = 0, nb = 0, ptr = priv_a;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        const uint32_t x = w[g];
        uint32_t e[4];
        e[0] = enc_lds32(lut_addr<0>(x, lanebase));
        e[1] = enc_lds32(lut_addr<1>(x, lanebase));
        e[2] = enc_lds32(lut_addr<2>(x, lanebase));
        e[3] = enc_lds32(lut_addr<3>(x, lanebase));
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t ej = g * 4 + j < nvalid ? e[j] : 0u;
            hi = __funnelshift_l(lo, hi, ej); lo = __funnelshift_l(ej, lo, ej);
            const uint32_t nb2 = nb + (ej & 31);
            if ((nb ^ nb2) & 32) { enc_sts32(ptr, __funnelshift_r(lo, hi, nb2)); ptr += 4; }
            nb = nb2;
        }
    }
    const uint32_t v = nb & 31;
    enc_sts32(ptr, shl_c(lo, 32 - v));
    if (v) enc_sts32(ptr + 4, 0);
    return nb;
}

// 16 symbols (4 little-endian words) starting at q; nvalid of them exist
__device__ __forceinline__ void load_syms16(const uint8_t* q, int nvalid, uint32_t w[4]) {
    if (nvalid == 16 && (reinterpret_cast<uintptr_t>(q) & 15) == 0) {
        const uint4 v = ld_stream_u4(reinterpret_cast<const uint4*>(q));
        w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            uint32_t v = 0;
#pragma unroll
            for (int b = 0; b < 4; ++b)
                if (i * 4 + b < nvalid) v |= (uint32_t)q[i * 4 + b] << (8 * b);
            w[i] = v;
        }
    }
}

// phase A, wide path: up to 8 symbols, codes of up to 32 bits
__device__ __forceinline__ uint32_t encode_thread_wide(const uint8_t* q, int nvalid, const uint2* lut64, uint32_t priv_a) {
    uint64_t acc = 0;
    uint32_t v = 0, nbits = 0, ptr = priv_a;
    for (int i = 0; i < nvalid; ++i) {
        const uint2 e = lut64[q[i]];                          // {length, right-aligned code}
        acc = (acc << e.x) | e.y;
        v += e.x; nbits += e.x;
        if (v >= 32) { v -= 32; enc_sts32(ptr, (uint32_t)(acc >> v)); ptr += 4; }
    }
    enc_sts32(ptr, shl_c((uint32_t)acc, 32 - v));
    if (v) enc_sts32(ptr + 4, 0);
    return nbits;
}

// ---------------------------------------------------------------------------------------------
// grid: K * cpc CTAs; CTA b serves chunk b / cpc, its group g the `mult` consecutive histogram
// segments starting at ((b % cpc) * ENC_GROUPS + g) * mult of that chunk (the range's first bit
// offset is seg_bitoff of its first segment)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(ENC_CTA, 2)
encode_kernel(const uint8_t* __restrict__ in, uint64_t n, uint32_t chunk_bytes, uint32_t spc, uint32_t cpc, uint32_t mult,
              const uint8_t* __restrict__ len_tab, const uint32_t* __restrict__ code_tab,
              const uint64_t* __restrict__ comp_off, const uint64_t* __restrict__ seg_bitoff,
              uint32_t K, uint8_t* __restrict__ out, uint64_t out_cap, uint32_t ident_on, int* status,
              const uint64_t* __restrict__ chain_ready, const uint32_t* __restrict__ comp_size) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const uint32_t t = threadIdx.x, lane = t & 31;
    const uint32_t grp = t / HZ_THREADS, tg = t % HZ_THREADS, wid = tg >> 5;
    const uint32_t k = blockIdx.x / cpc;
    const uint32_t s = ((blockIdx.x - k * cpc) * ENC_GROUPS + grp) * mult;   // first histogram segment of this group's range
    const uint32_t seg = k * spc + s;
    const uint32_t seg_bytes = mult * HZ_SEG_BYTES;                    // this group's range: `mult` histogram segments

    // geometry
    const uint64_t cbeg = (uint64_t)k * chunk_bytes;
    const uint64_t clen = n - cbeg < chunk_bytes ? n - cbeg : chunk_bytes;
    const uint64_t sbeg = (uint64_t)s * HZ_SEG_BYTES;
    const bool idle = s >= spc || sbeg >= clen;              // group-uniform
    const uint32_t slen = idle ? 0u : (uint32_t)(clen - sbeg < seg_bytes ? clen - sbeg : seg_bytes);
    const bool last_seg = sbeg + slen >= clen;
    const uint8_t* p = in + cbeg + sbeg;
    const uint8_t* q = p + tg * ENC_SPT;                     // this thread's symbols of tile 0
    const int lmode = (reinterpret_cast<uintptr_t>(q) & 31) == 0 ? 0 : ((reinterpret_cast<uintptr_t>(q) & 3) == 0 ? 1 : 2);
    const uint32_t full_tiles = slen / ENC_TILE;
    uint32_t w[8], wn[8];
    if (full_tiles) load_syms32(q, lmode, 32, wn);           // (the input does not depend on the codebook: in flight during the wait)
    // Chained launch (hz_codebook.cu: hist_chain_kernel): this kernel runs beside the tail of its predecessor, which raises
    // ready[k] once chunk k's lengths, codes, segment offsets and output offset are in memory.  They were written by a grid
    // that is still running, so they are read through L2 (ld.global.cg) - never through the non-coherent path, whose lines
    // (comp_off and seg_bitoff share theirs with neighbouring chunks) could predate the writes.  The wait is bounded.
    if (chain_ready) {
        if (t == 0) {
            uint64_t v; uint32_t spins = 0, ns = 64;
            for (;;) {
                asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(chain_ready + k) : "memory");
                if (v) break;
                if (++spins > (1u << 22)) { hz_set_status(status, HZ_ERR_CUDA); break; }
                __nanosleep(ns); if (ns < 1024) ns += ns;
            }
        }
        __syncthreads();
    }
    // every start-up load is issued before the first use (they are independent)
    const uint32_t mylen = __ldcg(len_tab + (size_t)k * 256 + tg);
    const uint32_t mycode = __ldcg(code_tab + (size_t)k * 256 + tg);
    const uint64_t chunk_off = __ldcg(comp_off + k);
    // the payload must fit the caller's buffer: the whole stream's (comp_off[K]) or, chained, this chunk's end
    const uint64_t need_bytes = chain_ready ? chunk_off + __ldcg(comp_size + k) : __ldcg(comp_off + K);
    const uint64_t seg_off = idle ? 0 : __ldcg(seg_bitoff + seg);
    if (need_bytes > out_cap) { if (t == 0) hz_set_status(status, HZ_ERR_OUT_TOO_SMALL); return; }
    // Identity chunk: all 256 symbols have 8-bit codes, so the canonical code of a symbol is the symbol itself
    // (CanonicalHuffman.java:99-132) and the chunk's bitstream is its plaintext - a byte copy (incompressible data).
    if (__syncthreads_and(ident_on && mylen == 8)) {
        if (!idle) hz_group_copy(out + chunk_off + sbeg, p, slen, tg, HZ_THREADS);
        return;
    }

    uint32_t* lut = reinterpret_cast<uint32_t*>(smem_raw);
    uint8_t* gsm = smem_raw + ENC_LUT_BYTES + grp * ENC_G_BYTES;
    uint32_t* dense = reinterpret_cast<uint32_t*>(gsm + ENC_G_DENSE);
    uint32_t* warp_tot = reinterpret_cast<uint32_t*>(gsm + ENC_G_WTOT);
    const uint32_t base_a = enc_smem_u32(smem_raw);
    const uint32_t dense_a = enc_pin(base_a + ENC_LUT_BYTES + grp * ENC_G_BYTES + ENC_G_DENSE);
    const uint32_t priv_a = enc_pin(base_a + ENC_LUT_BYTES + grp * ENC_G_BYTES + ENC_G_PRIV + tg * (ENC_PRIV_STRIDE * 4));
    const uint32_t bar_id = grp + 1;
    uint32_t par = 0;

    // codebook of this chunk (both groups fill half of every LUT row)
    const bool wide = __syncthreads_or(mylen > 27);    // block-uniform: codes longer than 27 bits
    const bool medium = __syncthreads_or(mylen > 16) && !wide;
    uint2* lut64 = reinterpret_cast<uint2*>(lut);     // wide: [sym] = {length, right-aligned code}
    if (!wide) {
        const uint32_t e = mylen ? (mycode << (32 - mylen)) | mylen : 0u;
        uint4 e4 = make_uint4(e, e, e, e);
        uint4* row = reinterpret_cast<uint4*>(&lut[tg * 32]) + grp * (8 / ENC_GROUPS);
#pragma unroll
        for (int i = 0; i < 8 / ENC_GROUPS; ++i) row[(i + lane) & (8 / ENC_GROUPS - 1)] = e4;   // rotated: fewer bank conflicts
    } else if (grp == 0) {
        lut64[tg] = make_uint2(mylen, mycode);
    }
    for (uint32_t i = tg; i < ENC_DENSE_WORDS / 4; i += HZ_THREADS) reinterpret_cast<uint4*>(dense)[i] = make_uint4(0, 0, 0, 0);
    __syncthreads();
    if (idle) return;                                  // from here on only group barriers

    // absolute bit address of the segment's first bit, and the 16-byte unit it falls in
    const uint64_t P0 = chunk_off * 8 + seg_off;
    const uint64_t out_addr = reinterpret_cast<uint64_t>(out);
    const uint64_t G0 = out_addr * 8 + P0;
    SegOut O;
    O.gptr = reinterpret_cast<uint4*>((G0 >> 7) << 4);
    O.cur = (uint32_t)(G0 & 127);
    O.skip = O.cur >> 3;                                // bytes of the first unit before this segment's first byte

    // the leading shared byte: previous segment's last (P0 & 7) bits, from its last 7 symbols
    if (tg == 0) {
        const uint32_t r = (uint32_t)(P0 & 7);
        if (r) {
            uint32_t bits = 0, have = 0;
            for (int j = 1; j <= 7 && have < r; ++j) {
                uint32_t sym = p[-j];
                uint32_t l = __ldcg(len_tab + (size_t)k * 256 + sym);
                uint32_t c = __ldcg(code_tab + (size_t)k * 256 + sym);
                uint32_t take = l < 8 ? l : 8;
                bits |= (c & ((1u << take) - 1)) << have;
                have += take;
            }
            bits &= (1u << r) - 1;
            const uint32_t b0 = O.cur - r;                       // r bits inside one byte -> one word
            enc_or32(dense_a + ((b0 >> 5) << 2), bits << (32 - (b0 & 31) - r));
        }
    }

    if (medium) {
        // ---- medium path: codes of 17..27 bits, 16 symbols per thread per tile --------------------
        const uint32_t lanebase = enc_pin(base_a + (lane << 2));
        uint32_t x[4], xn[4];
        auto nvalid_of = [&](uint32_t tile) -> int {
            const uint32_t first = tile + tg * ENC_MED_SPT;
            return first >= slen ? 0 : (slen - first >= ENC_MED_SPT ? ENC_MED_SPT : (int)(slen - first));
        };
        load_syms16(p + tg * ENC_MED_SPT, nvalid_of(0), xn);
        for (uint32_t tile = 0; tile < slen; tile += ENC_MED_TILE) {
#pragma unroll
            for (int j = 0; j < 4; ++j) x[j] = xn[j];
            if (tile + ENC_MED_TILE < slen) load_syms16(p + tile + ENC_MED_TILE + tg * ENC_MED_SPT, nvalid_of(tile + ENC_MED_TILE), xn);
            const uint32_t nbits = encode_thread_med(x, nvalid_of(tile), lanebase, priv_a);
            uint32_t tile_bits;
            const uint32_t off = O.cur + group_excl_scan(nbits, warp_tot, par, bar_id, lane, wid, &tile_bits);
            par ^= 1;
            merge_or(priv_a, dense_a, off, nbits);
            group_sync(bar_id);
            O.cur += tile_bits;
            flush_tile(dense, O, tg);
        }
    } else if (!wide) {
        // ---- fast path: 32 symbols per thread per tile, software-pipelined loads ------------------
        const uint32_t lanebase = enc_pin(base_a + (lane << 2));
        // one full tile: A (codewords -> private word-aligned bit string), S (exclusive scan of the
        // per-thread bit counts: one barrier, which also orders F of the previous tile before B of
        // this one), B (private -> dense), barrier, F (dense -> global)
        auto full_tile = [&](const uint32_t (&x)[8]) {
            const uint32_t nbits = encode_thread32<false>(x, 32, lanebase, priv_a);
            uint32_t tile_bits;
            const uint32_t off = O.cur + group_excl_scan(nbits, warp_tot, par, bar_id, lane, wid, &tile_bits);
            par ^= 1;
            merge_shuffle(priv_a, dense_a, off, nbits, lane);
            group_sync(bar_id);
            O.cur += tile_bits;
            flush_tile(dense, O, tg);
        };
        // The prefetched words are copied (8 FMA-pipe moves) before the next load is issued: the
        // copies wait for the previous load, so the new load's scoreboard is not waited on until
        // the next tile (ping-ponging two register sets made every other tile stall on DRAM).
        for (uint32_t i = 0; i < full_tiles; ++i) {
#pragma unroll
            for (int j = 0; j < 8; ++j) w[j] = wn[j];
            q += ENC_TILE;
            if (i + 1 < full_tiles) load_syms32(q, lmode, 32, wn);
            full_tile(w);
        }
        const uint32_t done = full_tiles * ENC_TILE;
        if (done < slen) {                                  // ragged last tile (group-uniform)
            const uint32_t first = done + tg * ENC_SPT;
            const int nvalid = first >= slen ? 0 : (slen - first >= ENC_SPT ? ENC_SPT : (int)(slen - first));
            load_syms32(q, nvalid == 32 ? lmode : 2, nvalid, w);
            const uint32_t nbits = encode_thread32<true>(w, nvalid, lanebase, priv_a);
            uint32_t tile_bits;
            const uint32_t off = O.cur + group_excl_scan(nbits, warp_tot, par, bar_id, lane, wid, &tile_bits);
            par ^= 1;
            merge_or(priv_a, dense_a, off, nbits);
            group_sync(bar_id);
            O.cur += tile_bits;
            flush_tile(dense, O, tg);
        }
    } else {
        // ---- wide path: codes of up to 32 bits, 8 symbols per thread per tile --------------------
        for (uint32_t tile = 0; tile < slen; tile += ENC_WIDE_TILE) {
            const uint32_t first = tile + tg * ENC_WIDE_SPT;
            const int nvalid = first >= slen ? 0 : (slen - first >= ENC_WIDE_SPT ? ENC_WIDE_SPT : (int)(slen - first));
            const uint32_t nbits = encode_thread_wide(p + first, nvalid, lut64, priv_a);
            uint32_t tile_bits;
            const uint32_t off = O.cur + group_excl_scan(nbits, warp_tot, par, bar_id, lane, wid, &tile_bits);
            par ^= 1;
            merge_or(priv_a, dense_a, off, nbits);
            group_sync(bar_id);
            O.cur += tile_bits;
            flush_tile(dense, O, tg);
        }
    }
    group_sync(bar_id);
    flush_tail(dense, O, last_seg, tg);
}

int hzk_encode(hz_ctx* ctx, const uint8_t* d_in, uint64_t n, uint32_t chunk_bytes, uint32_t K,
               const uint8_t* d_len, const uint32_t* d_code, const uint64_t* d_comp_off,
               const uint64_t* d_seg_bitoff, uint8_t* d_out, uint64_t out_cap,
               const uint64_t* d_chain_ready, const uint32_t* d_comp_size) {
    if (K == 0) return HZ_OK;
    const uint32_t spc = (chunk_bytes + HZ_SEG_BYTES - 1) / HZ_SEG_BYTES;
    // longer ranges amortise the per-range start-up (LUT build, zeroing, first loads); short
    // chunks keep them short so that both groups of a CTA have work
    const uint32_t mult = hz_range_mult(spc, ctx->knobs.range_mult);
    const uint32_t rpc = (spc + mult - 1) / mult;                      // ranges per chunk
    const uint32_t cpc = (rpc + ENC_GROUPS - 1) / ENC_GROUPS;
    const uint64_t grid = (uint64_t)K * cpc;
    if (grid > 0x7fffffffull) return hz_fail(ctx, HZ_ERR_ARG, "too many segments");
    if (!ctx->attr_encode) {
        HZ_CUDA(ctx, cudaFuncSetAttribute(encode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ENC_SMEM_BYTES));
        ctx->attr_encode = true;
    }
    const uint32_t ident_on = ctx->knobs.ident;   // developer knob: HZ_IDENT=0 sends identity chunks through the bit packer
    if (!d_chain_ready) {
        HZ_LAUNCH(ctx, "encode", encode_kernel, (unsigned)grid, ENC_CTA, ENC_SMEM_BYTES,
                  d_in, n, chunk_bytes, spc, cpc, mult, d_len, d_code, d_comp_off, d_seg_bitoff, K, d_out, out_cap, ident_on, ctx->d_status,
                  (const uint64_t*)nullptr, (const uint32_t*)nullptr);
        return HZ_OK;
    }
    // chained: programmatic stream serialization - the grid may start once every CTA of hist_chain_kernel has issued
    // griddepcontrol.launch_dependents; it never executes griddepcontrol.wait, the per-chunk ready flags order the data.
    // (With the per-kernel profiler on, the event between the two launches serialises them: same results, no overlap.)
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(ENC_CTA); cfg.dynamicSmemBytes = ENC_SMEM_BYTES; cfg.stream = ctx->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    hz_prof_begin(ctx);
    cudaError_t e = cudaLaunchKernelEx(&cfg, encode_kernel, d_in, n, chunk_bytes, spc, cpc, mult, d_len, d_code, d_comp_off, d_seg_bitoff,
                                       K, d_out, out_cap, ident_on, ctx->d_status, d_chain_ready, d_comp_size);
    ctx->launches++;
    hz_prof_end(ctx, "encode");
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) return hz_cuda_fail(ctx, e, "encode (chained)");
    return HZ_OK;
}
