// hz_decode.cu — chunked parallel Huffman decode.
//
// Replaces CanonicalHuffman.generateCanonicalCodesFromLengths + TableBasedHuffmanDecoder.decode
// (core/CanonicalHuffman.java:141-146, core/TableBasedHuffmanDecoder.java:36-152, driven by
// CpuCompressionService.decodeChunkParallel, service/cpu/CpuCompressionService.java:511-532).
// A chunk of the .dcz payload is ONE sequential bitstream without restart markers, and the
// container must stay bit-identical, so parallelism inside a chunk comes from the
// self-synchronisation property of Huffman codes.  The stream of a chunk is cut into
// subsequences of DEC_SUB_BITS bits (one per thread); 256 subsequences form a sequence (one CTA
// pass); a CTA owns DEC_SEQ_PER_CTA consecutive sequences.
//
//   plan    (1 CTA)       per-chunk subsequence / sequence / CTA counts and their prefix sums.
//   tables  (1 CTA/chunk) multi-symbol lookup tables of every chunk that spans several CTAs,
//                         built once into global memory (single-CTA chunks build them in place).
//   sync    (many CTAs)   the compressed bytes of a sequence are staged in shared memory with a
//                         1-D TMA bulk copy (cp.async.bulk + mbarrier, double buffered).  Thread i
//                         starts DEC_OVERLAP_BITS before subsequence i (a guess), records where it
//                         crosses INTO the subsequence (entry), keeps decoding and counting
//                         codewords until it crosses OUT (exit).  The chain is valid when
//                         exit[i-1] == entry[i]; mismatches inside a CTA are repaired by
//                         re-decoding from the neighbour's exit; the CTA's first subsequence is
//                         left to:
//   fix     (1 CTA/chunk) compares every CTA boundary, re-walks from the true position where the
//                         guess was wrong, then scans sequence symbol counts into output offsets.
//   write   (many CTAs)   decodes every subsequence again from its verified entry, 1-3 symbols per
//                         table lookup, into a per-warp shared-memory window that is written out
//                         with aligned 128-bit stores.
//
// Tables (LUTB = 12 bits of look-ahead):
//   slut u16: ltot | l0<<6 | n<<12   bits / number of all n (<=12) complete codewords inside the
//             12 bits and the first codeword's length (counting pass).  A prefix that starts a
//             code longer than 12 bits has n = 1 and its length when every code under the prefix
//             has the same length, else n = 0 with ltot = shortest, l0 = longest candidate.
//   wlut u32x2: .x = s0 | s1<<8 | s2<<16 | s3<<24 (up to four symbols per lookup), .y = ltot | 8n<<16:
//             ONE add of .y to the write pass's packed counter (stream bit position | output bits<<16)
//             advances both.  .y bit 31 alone: every code under the prefix has the same length l = .y & 63
//             (> LUTB) and its symbol is sorted[k + (the l - LUTB stream bits after the prefix)]: the
//             sorted-symbol array is the second level; .x = (DEC_W_SORTED_REL + k) << 5 | (32 + LUTB - l),
//             i.e. a wlut-relative shared address and the right-shift that isolates those bits, so the
//             write loop resolves such codes with five predicated instructions and no branch.
//             .y bits 31+30: several candidate lengths or no code, .x = shortest | longest<<6.
// Codes whose used lengths are all equal never self-synchronise but need no synchronisation
// either: entries are computed arithmetically.
#include <cstring>
#include "hz_decode_tables.cuh"

#ifndef DEC_SUB_WORDS
#define DEC_SUB_WORDS 17                      // odd: subsequences start in different smem banks
#endif
#define DEC_SUB_BITS (DEC_SUB_WORDS * 32)     // 544
#define DEC_SUB_BYTES (DEC_SUB_WORDS * 4)     // 68
#define DEC_SEQ_BYTES (DT * DEC_SUB_BYTES)    // 17408
#define DEC_SEQ_BITS (DT * DEC_SUB_BITS)
#ifndef DEC_OVERLAP_BITS
#define DEC_OVERLAP_BITS 128
#endif
#define DEC_OVERLAP_BYTES (DEC_OVERLAP_BITS / 8)
#ifndef DEC_SEQ_PER_CTA
#define DEC_SEQ_PER_CTA 12                    // divisible by 1, 2 and 3 write groups
#endif
#define DEC_SUBS_PER_CTA (DT * DEC_SEQ_PER_CTA)
// staged bytes per sequence: 16 alignment slack + 16 overlap + sequence + 32 look-ahead
#define DEC_STAGE_BYTES ((16 + DEC_OVERLAP_BYTES + DEC_SEQ_BYTES + 32 + 15) & ~15)
#define DEC_STAGE_WORDS (DEC_STAGE_BYTES / 4)
#define DEC_WIN_MIN 2304                      // per-warp output window of the write kernel (runtime sized)
#define DEC_WIN_MAX 9216
#define DEC_TAB_PREBUILD_CAP (6ull << 30)        // scratch the prebuilt tables of single-CTA chunks may take (41 KiB per chunk)



// ---------------------------------------------------------------------------------------------
// plan
// ---------------------------------------------------------------------------------------------
struct DecPlan {
    uint32_t* nsub; uint32_t* sub_base; uint32_t* seq_base; uint32_t* cta_base; uint32_t* tab_idx; uint64_t* orig_off;
    uint32_t* islice;          // [K+1] copy slices of the identity chunks before chunk k
};
#define DEC_IDENT_SLICE 65536u

__device__ __forceinline__ uint32_t ceil_div_u64(uint64_t a, uint32_t b) { return (uint32_t)((a + b - 1) / b); }

__global__ void __launch_bounds__(1024)
dec_plan_kernel(const uint64_t* __restrict__ comp_off, uint64_t comp_bytes, const uint32_t* __restrict__ comp_size,
                const uint32_t* __restrict__ orig_size, const uint64_t* __restrict__ orig_off_in, uint32_t K, DecPlan P,
                uint32_t tab_min_seq, uint8_t* __restrict__ ident, int* status) {
    __shared__ uint64_t part[6][1024];
    const uint32_t t = threadIdx.x;
    const uint32_t per = (K + 1023) / 1024;
    const uint32_t lo = min(K, t * per), hi = min(K, lo + per);
    uint64_t s0 = 0, s1 = 0, s2 = 0, s3 = 0, s4 = 0, s5 = 0;
    // a chunk must lie inside the addressable stream (untrusted footer fields reach this ABI): one that does not is
    // skipped (no subsequences, no copy slices) and the call reports HZ_ERR_ARG
    for (uint32_t i = lo; i < hi; ++i)
        if (!(comp_off[i] <= comp_bytes && comp_size[i] <= comp_bytes - comp_off[i])) { ident[i] = 2; hz_set_status(status, HZ_ERR_ARG); }
    for (uint32_t i = lo; i < hi; ++i) {
        s5 += ident[i] == 1 ? (orig_size[i] + DEC_IDENT_SLICE - 1) / DEC_IDENT_SLICE : 0u;
        uint32_t ns = (orig_size[i] && !ident[i]) ? max(1u, ceil_div_u64((uint64_t)comp_size[i] * 8, DEC_SUB_BITS)) : 0;
        uint32_t nq = (ns + DT - 1) / DT;
        s0 += ns; s1 += nq; s2 += (nq + DEC_SEQ_PER_CTA - 1) / DEC_SEQ_PER_CTA; s3 += orig_size[i];
        s4 += nq > tab_min_seq;
    }
    part[0][t] = s0; part[1][t] = s1; part[2][t] = s2; part[3][t] = s3; part[4][t] = s4; part[5][t] = s5;
    __syncthreads();
    if (t < 6) {
        uint64_t a = 0;
        for (int j = 0; j < 1024; ++j) { uint64_t x = part[t][j]; part[t][j] = a; a += x; }
        if (t == 0) P.sub_base[K] = (uint32_t)a;
        if (t == 1) P.seq_base[K] = (uint32_t)a;
        if (t == 2) P.cta_base[K] = (uint32_t)a;
        if (t == 3) P.orig_off[K] = a;
        if (t == 5) P.islice[K] = (uint32_t)a;
    }
    __syncthreads();
    s0 = part[0][t]; s1 = part[1][t]; s2 = part[2][t]; s3 = part[3][t]; s4 = part[4][t]; s5 = part[5][t];
    for (uint32_t i = lo; i < hi; ++i) {
        P.islice[i] = (uint32_t)s5;
        s5 += ident[i] == 1 ? (orig_size[i] + DEC_IDENT_SLICE - 1) / DEC_IDENT_SLICE : 0u;
        uint32_t ns = (orig_size[i] && !ident[i]) ? max(1u, ceil_div_u64((uint64_t)comp_size[i] * 8, DEC_SUB_BITS)) : 0;
        uint32_t nq = (ns + DT - 1) / DT;
        P.nsub[i] = ns; P.sub_base[i] = (uint32_t)s0; P.seq_base[i] = (uint32_t)s1; P.cta_base[i] = (uint32_t)s2;
        P.orig_off[i] = orig_off_in ? orig_off_in[i] : s3;
        P.tab_idx[i] = nq > tab_min_seq ? (uint32_t)s4 : DEC_NO_TABLE;
        s0 += ns; s1 += nq; s2 += (nq + DEC_SEQ_PER_CTA - 1) / DEC_SEQ_PER_CTA; s3 += orig_size[i];
        s4 += nq > tab_min_seq;
    }
}

// ---------------------------------------------------------------------------------------------
// Identity chunks.  When all 256 symbols of a chunk have 8-bit codes (incompressible data: every byte value
// about equally frequent), generateCanonicalCodesFromLengths assigns code[s] == s (CanonicalHuffman.java:
// 99-132: codes of one length are consecutive in symbol order), so the chunk's payload IS its plaintext and
// decoding it is a byte copy at HBM speed instead of a table walk.  The flags kernel marks such chunks (one
// warp per chunk); the plan gives them no subsequences, so the sync / fix / write kernels never see them;
// the copy kernel (persistent grid over 64 KiB slices, located through the plan's slice prefix sums) moves them.  Bytes the stream lacks
// (comp_size < orig_size) decode as symbol 0, like zero bits past the end (TableBasedHuffmanDecoder.java:204-208).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(DT)
dec_ident_flags_kernel(const uint8_t* __restrict__ len_tab, const uint32_t* __restrict__ orig_size, uint32_t K,
                       uint8_t* __restrict__ ident, uint32_t enable) {
    const uint32_t k = blockIdx.x * (DT / 32) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (k >= K) return;
    const uint2 v = reinterpret_cast<const uint2*>(len_tab + (size_t)k * 256)[lane];
    const bool all8 = __all_sync(0xffffffffu, v.x == 0x08080808u && v.y == 0x08080808u);
    if (lane == 0) ident[k] = enable && all8 && orig_size[k] != 0;
}

__global__ void __launch_bounds__(DT)
dec_ident_copy_kernel(const uint8_t* __restrict__ comp, const uint64_t* __restrict__ comp_off,
                      const uint32_t* __restrict__ comp_size, const uint32_t* __restrict__ orig_size,
                      const uint64_t* __restrict__ orig_off, const uint32_t* __restrict__ islice, uint32_t K,
                      uint8_t* __restrict__ out, uint64_t out_cap, int* status) {
    const uint32_t total = islice[K];
    for (uint32_t j = blockIdx.x; j < total; j += gridDim.x) {
        uint32_t lo = 0, hi = K;                       // largest k with islice[k] <= j (identity chunks have >= 1 slice)
        while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) >> 1;
            if (islice[mid] <= j) lo = mid; else hi = mid;
        }
        const uint32_t k = lo;
        const uint32_t osize = orig_size[k], csize = comp_size[k];
        const uint64_t ooff = orig_off[k];
        if (ooff + osize > out_cap) { if (threadIdx.x == 0) hz_set_status(status, HZ_ERR_OUT_TOO_SMALL); return; }
        const uint32_t have = csize < osize ? csize : osize;
        const uint32_t b0 = (j - islice[k]) * DEC_IDENT_SLICE;
        const uint32_t b1 = min(osize, b0 + DEC_IDENT_SLICE);
        const uint32_t c1 = min(have, b1);
        if (c1 > b0) hz_group_copy(out + ooff + b0, comp + comp_off[k] + b0, c1 - b0, threadIdx.x, DT);
        for (uint32_t b = max(b0, c1) + threadIdx.x; b < b1; b += DT) out[ooff + b] = 0;
    }
}

// chunk that owns global CTA index `b` (largest k with cta_base[k] <= b)
__device__ __forceinline__ uint32_t find_chunk(const uint32_t* __restrict__ cta_base, uint32_t K, uint32_t b) {
    uint32_t lo = 0, hi = K;          // invariant: cta_base[lo] <= b < cta_base[hi]
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (cta_base[mid] <= b) lo = mid; else hi = mid;
    }
    return lo;
}

// ---------------------------------------------------------------------------------------------
// tables kernel: one CTA per chunk; chunks handled by a single CTA build their tables in place
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(DT)
dec_tables_kernel(const uint8_t* __restrict__ len_tab, DecPlan P, uint8_t* __restrict__ tables, int* status) {
    __shared__ __align__(16) uint16_t slut[LUTN];
    __shared__ __align__(16) uint8_t scratch[DEC_BUILD_SCRATCH];
    __shared__ __align__(16) uint8_t aux_raw[1024];
    const uint32_t k = blockIdx.x;
    const uint32_t ti = P.tab_idx[k];
    if (ti == DEC_NO_TABLE) return;
    DecAux& A = *reinterpret_cast<DecAux*>(aux_raw);
    uint8_t* dst = tables + (size_t)ti * DEC_TABLE_BYTES;
    build_tables<true, true>(A, reinterpret_cast<uint2*>(dst + DEC_TAB_W), slut, scratch, len_tab + (size_t)k * 256);
    if (A.bad && threadIdx.x == 0) hz_set_status(status, HZ_ERR_BAD_LENGTHS);
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < LUTN * 2 / 16; i += DT) reinterpret_cast<uint4*>(dst + DEC_TAB_S)[i] = reinterpret_cast<uint4*>(slut)[i];
    for (uint32_t i = threadIdx.x; i < 1024 / 16; i += DT) reinterpret_cast<uint4*>(dst + DEC_TAB_AUX)[i] = reinterpret_cast<uint4*>(aux_raw)[i];
}

// ---------------------------------------------------------------------------------------------
// staging of one sequence's compressed bytes: 1-D TMA bulk copy + mbarrier
// ---------------------------------------------------------------------------------------------

struct StageGeom {
    uint64_t a0;         // 16-byte aligned global address of stage byte 0
    int64_t vlo, vhi;    // stage-relative byte range that belongs to the chunk
    int64_t tlo, thi;    // stage-relative byte range delivered by the bulk copy
    uint32_t bit0;       // stage-relative bit index of the sequence's first bit
};

__device__ __forceinline__ StageGeom stage_geom(const uint8_t* comp, uint64_t comp_bytes, uint64_t chunk_off,
                                                uint32_t chunk_size, uint32_t sq) {
    StageGeom g;
    const uint64_t cb = reinterpret_cast<uint64_t>(comp) + chunk_off;
    const uint64_t seq0 = cb + (uint64_t)sq * DEC_SEQ_BYTES;                 // address of the sequence's first byte
    const uint64_t lo = seq0 - DEC_OVERLAP_BYTES;
    g.a0 = lo & ~(uint64_t)15;
    g.bit0 = (uint32_t)(seq0 - g.a0) * 8;
    g.vlo = (int64_t)cb - (int64_t)g.a0;
    g.vhi = g.vlo + chunk_size;
    if (g.vlo < 0) g.vlo = 0;
    if (g.vhi > DEC_STAGE_BYTES) g.vhi = DEC_STAGE_BYTES;
    if (g.vhi < g.vlo) g.vhi = g.vlo;
    const uint64_t blo = (reinterpret_cast<uint64_t>(comp) + 15) & ~(uint64_t)15;
    const uint64_t bhi = (reinterpret_cast<uint64_t>(comp) + comp_bytes) & ~(uint64_t)15;
    uint64_t tl = g.a0 > blo ? g.a0 : blo;
    uint64_t th = g.a0 + DEC_STAGE_BYTES < bhi ? g.a0 + DEC_STAGE_BYTES : bhi;
    if (th < tl) th = tl;
    g.tlo = (int64_t)(tl - g.a0); g.thi = (int64_t)(th - g.a0);
    return g;
}

// issued by ONE thread
__device__ __forceinline__ void stage_issue(uint8_t* stage, uint64_t* bar, const StageGeom& g) {
    const uint32_t bytes = (uint32_t)(g.thi - g.tlo);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (bytes) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(stage + g.tlo)), "l"(g.a0 + (uint64_t)g.tlo), "r"(bytes), "r"(smem_u32(bar)) : "memory");
    } else {
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
    }
}

// after the bulk copy landed: bytes outside the chunk read as zero (TableBasedHuffmanDecoder.java:204-208),
// chunk bytes the 16-byte aligned copy could not deliver are fetched one by one.  All DT threads.
__device__ __forceinline__ void stage_fixup(uint8_t* stage, const StageGeom& g, uint32_t tg) {
    const int64_t glo = g.vlo > g.tlo ? g.vlo : g.tlo, ghi = g.vhi < g.thi ? g.vhi : g.thi;   // good bytes
    if (glo == 0 && ghi == DEC_STAGE_BYTES) return;
    for (int64_t u = tg; u < DEC_STAGE_BYTES / 16; u += DT) {
        const int64_t b0 = u * 16;
        if (b0 >= glo && b0 + 16 <= ghi) continue;
        for (int64_t b = b0; b < b0 + 16; ++b) {
            if (b >= glo && b < ghi) continue;
            uint8_t v = 0;
            if (b >= g.vlo && b < g.vhi) v = *reinterpret_cast<const uint8_t*>(g.a0 + (uint64_t)b);
            stage[b] = v;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// bit reader over the staged bytes: (hi, lo) = 64 upcoming stream bits, sh = consumed bits of hi
// ---------------------------------------------------------------------------------------------

struct BitRd {
    uint32_t wa;           // shared address of the word held in `lo`
    uint32_t hi, lo;       // the (byte-swapped) stage words holding bits [32*(pos>>5), +64)
};
__device__ __forceinline__ void rd_seek(BitRd& r, uint32_t stage_a, uint32_t pos) {
    r.wa = stage_a + ((pos >> 5) << 2) + 4;
    r.hi = bswap32(lds32(r.wa - 4)); r.lo = bswap32(lds32(r.wa));
}
// the 32 stream bits that start at pos (the funnel shift takes pos mod 32)
__device__ __forceinline__ uint32_t rd_peek32(const BitRd& r, uint32_t pos) { return __funnelshift_l(r.lo, r.hi, pos); }
__device__ __forceinline__ void rd_skip(BitRd& r, uint32_t& pos, uint32_t l) {       // l <= 32
    const uint32_t np = pos + l;
    if ((np ^ pos) >= 32) { r.hi = r.lo; r.wa += 4; r.lo = bswap32(lds32(r.wa)); }
    pos = np;
}

struct SyncSmem {
    __align__(16) uint8_t stage[2][DEC_STAGE_BYTES];
    __align__(16) uint16_t slut[LUTN];
    __align__(16) uint8_t aux[1024];
    __align__(8) uint64_t bar[2];
    uint32_t s_exit[DT];
    uint32_t s_red[DT / 32];
    uint32_t s_k;
};


// Right shifts inside the hand-written loops stay SHF: taking them as hi32(x * 2^(32 - s)) on the FMA pipe was
// measured on B200 and rejected (dec_sync 1.93 -> 2.12 ms, dec_write 3.66 -> 3.71 ms).
#define HZ_SHR20(D, S) "shr.u32 " D ", " S ", 20;\n"
#define HZ_SHR16(D, S) "shr.u32 " D ", " S ", 16;\n"
#define HZ_SHR12(D, S) "shr.u32 " D ", " S ", 12;\n"
#define HZ_ADD4(PRED, R, TWO) PRED " add.u32 " R ", " R ", 4;\n"
#define HZ_W_TWO_OPERAND

// Advance from `pos` to the first codeword boundary >= limit; returns the number of codewords
// that began before `limit`.  Unmatched patterns consume one bit.
__device__ __forceinline__ uint32_t advance(const DecAux& A, uint32_t slut, BitRd& r, uint32_t& pos, uint32_t limit) {
    uint32_t cnt = 0;
    if (pos + LUTB <= limit) {
        // Main loop (hand-written PTX, two lookups per trip so that the position ping-pongs between two
        // registers): all complete codewords of the 12-bit window per lookup.  The stage word after `lo` is
        // loaded one refill ahead (raw; byte-swapped when it becomes `lo`), the refill is four predicated
        // instructions.  Entries with n == 0 (several candidate lengths / no code) branch to an out-of-line
        // search and come back into the same iteration, so the warp reconverges every lookup.
        uint32_t nx = lds32(r.wa + 4);
        const uint32_t lim12 = limit - LUTB;
        // 2 that the compiler cannot fold (grids are one-dimensional): the table address is then an IMAD (FMA pipe)
        // instead of an LEA on the integer ALU pipe, the busiest pipe of this loop
        const uint32_t two = 2u * gridDim.y;
#define HZ_ASTEP(PI, PO, SFX)                                                  \
    "shf.l.wrap.b32 v, %2, %1, " PI ";\n"                                      \
    HZ_SHR20("ix", "v")                                                        \
    "mad.lo.u32 ix, ix, %9, %6;\n"                                             \
    "ld.shared.u16 e, [ix];\n"                                                 \
    "and.b32 l, e, 63;\n"                                                      \
    HZ_SHR12("n", "e")                                                         \
    "setp.lt.u32 pz, e, 4096;\n"                                               \
    "@pz bra HZA_RARE" SFX ";\n"                                               \
    "HZA_BACK" SFX ":\n"                                                       \
    "add.u32 " PO ", " PI ", l;\n"                                             \
    "add.u32 %5, %5, n;\n"                                                     \
    "xor.b32 t, " PI ", " PO ";\n"                                             \
    "and.b32 t, t, 32;\n"                                                      \
    "setp.ne.u32 p0, t, 0;\n"                                                  \
    "@p0 mov.u32 %1, %2;\n"                                                    \
    "@p0 prmt.b32 %2, %4, z, 0x0123;\n"                                        \
    "@p0 ld.shared.u32 %4, [%3+8];\n"                                          \
    HZ_ADD4("@p0", "%3", "%9")                                                 \
    "setp.gt.u32 pc, " PO ", %7;\n"
#define HZ_ARARE(SFX)                                                          \
    "HZA_RARE" SFX ":\n"                                                       \
    "shr.u32 m, e, 6;\n"                                                       \
    "and.b32 m, m, 63;\n"                                                      \
    "mov.u32 auxb, %8;\n"                                                      \
    HZ_PTX_LONGLEN("A" SFX)                                                    \
    "max.u32 l, l, 1;\n"                                                       \
    "mov.u32 n, 1;\n"                                                          \
    "bra HZA_BACK" SFX ";\n"
        asm volatile(
            "{\n"
            ".reg .pred pz, p0, pc, pq;\n"
            ".reg .u32 Q, v, ix, e, l, n, m, t, u, a, z, auxb;\n"
            "mov.u32 z, 0;\n"
            "HZA_TOP:\n"
            HZ_ASTEP("%0", "Q", "1")
            "@pc bra HZA_ENDQ;\n"
            HZ_ASTEP("Q", "%0", "2")
            "@!pc bra HZA_TOP;\n"
            "bra HZA_END;\n"
            HZ_ARARE("1")
            HZ_ARARE("2")
            "HZA_ENDQ:\n"
            "mov.u32 %0, Q;\n"
            "HZA_END:\n"
            "}\n"
            : "+r"(pos), "+r"(r.hi), "+r"(r.lo), "+r"(r.wa), "+r"(nx), "+r"(cnt)
            : "r"(slut), "r"(lim12), "r"(slut + (uint32_t)(offsetof(SyncSmem, aux) - offsetof(SyncSmem, slut))), "r"(two)
            : "memory");
#undef HZ_ASTEP
#undef HZ_ARARE
    }
    while (pos < limit) {
        const uint32_t v = rd_peek32(r, pos);
        const uint32_t e = lds16(slut + 2 * (v >> (32 - LUTB)));
        uint32_t l = (e >> 6) & 63;
        if ((e >> 12) == 0) { l = long_len(A, v, e & 63, l); if (!l) l = 1; }
        rd_skip(r, pos, l); ++cnt;
    }
    return cnt;
}

// record layout: entry (8) | exit (8) | count (16)
__device__ __forceinline__ uint32_t pack_rec(uint32_t entry, uint32_t exitv, uint32_t count) {
    return entry | (exitv << 8) | (count << 16);
}


__global__ void __launch_bounds__(DT)
dec_sync_kernel(const uint8_t* __restrict__ comp, uint64_t comp_bytes, const uint64_t* __restrict__ comp_off,
                const uint32_t* __restrict__ comp_size, const uint8_t* __restrict__ len_tab,
                uint32_t K, DecPlan P, const uint8_t* __restrict__ tables,
                uint32_t* __restrict__ rec, uint32_t* __restrict__ seqcnt, int* status) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    SyncSmem& S = *reinterpret_cast<SyncSmem*>(smem_raw);
    const uint32_t t = threadIdx.x, lane = t & 31, wid = t >> 5;
    if (blockIdx.x >= P.cta_base[K]) return;
    if (t == 0) {
        S.s_k = find_chunk(P.cta_base, K, blockIdx.x);
        mbar_init(&S.bar[0], 1); mbar_init(&S.bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const uint32_t k = S.s_k;
    DecAux& A = *reinterpret_cast<DecAux*>(S.aux);
    const uint32_t ti = P.tab_idx[k];
    if (ti == DEC_NO_TABLE) {
        build_tables<false, true>(A, nullptr, S.slut, S.stage[0], len_tab + (size_t)k * 256);
    } else {
        const uint8_t* tb = tables + (size_t)ti * DEC_TABLE_BYTES;
        copy_g2s16(S.slut, tb + DEC_TAB_S, LUTN * 2);
        copy_g2s16(S.aux, tb + DEC_TAB_AUX, 1024);
    }
    __syncthreads();
    if (A.bad) { if (t == 0) hz_set_status(status, HZ_ERR_BAD_LENGTHS); return; }

    const uint64_t coff = comp_off[k];
    const uint32_t csize = comp_size[k];
    const uint32_t nsub = P.nsub[k];
    const uint32_t nseq = (nsub + DT - 1) / DT;
    const uint32_t cta_in_chunk = blockIdx.x - P.cta_base[k];
    const uint32_t sq0 = cta_in_chunk * DEC_SEQ_PER_CTA;
    const uint32_t nq = min((uint32_t)DEC_SEQ_PER_CTA, nseq - sq0);
    const uint32_t U = (uint32_t)A.uniform;

    if (t == 0) stage_issue(S.stage[0], &S.bar[0], stage_geom(comp, comp_bytes, coff, csize, sq0));
    uint32_t carry_exit = 0;         // exit of the previous sequence's last subsequence (q > 0)
    for (uint32_t q = 0; q < nq; ++q) {
        const uint32_t sq = sq0 + q;
        const uint32_t b = q & 1;
        if (t == 0 && q + 1 < nq)
            stage_issue(S.stage[b ^ 1], &S.bar[b ^ 1], stage_geom(comp, comp_bytes, coff, csize, sq + 1));
        const StageGeom g = stage_geom(comp, comp_bytes, coff, csize, sq);
        mbar_wait(&S.bar[b], (q >> 1) & 1);
        stage_fixup(S.stage[b], g, t);
        __syncthreads();

        BitRd r;
        const uint32_t stage_a = pin_reg(smem_u32(S.stage[b])), slut_a = pin_reg(smem_u32(S.slut));
        const uint32_t i = sq * DT + t;
        const bool active = i < nsub;
        const uint32_t nominal = g.bit0 + t * DEC_SUB_BITS;       // stage-relative
        const uint32_t end = nominal + DEC_SUB_BITS;
        uint32_t entry = 0, count = 0, exitv = 0;
        if (active) {
            uint32_t pos;
            if (U) {                     // equal-length code: boundaries are the multiples of U
                const uint64_t nomc = (uint64_t)i * DEC_SUB_BITS;
                entry = (uint32_t)((U - nomc % U) % U);
                // every U-bit pattern counts as one codeword (an unmatched pattern of a one-symbol code
                // consumes one bit = U), so count and exit follow from arithmetic alone
                count = (DEC_SUB_BITS - entry + U - 1) / U;
                exitv = entry + count * U - DEC_SUB_BITS;
            } else if (i == 0) {
                pos = nominal; rd_seek(r, stage_a, pos);
            } else if (t == 0 && q > 0) {
                entry = carry_exit; pos = nominal + entry; rd_seek(r, stage_a, pos);
            } else {
                pos = nominal - DEC_OVERLAP_BITS; rd_seek(r, stage_a, pos);
                advance(A, slut_a, r, pos, nominal);
                entry = pos - nominal;
            }
            if (!U) {
                count = advance(A, slut_a, r, pos, end);
                exitv = pos - end;
            }
        }
        S.s_exit[t] = exitv;
        __syncthreads();
        // repair mismatches inside the sequence until the chain is consistent
        if (!U) {
            for (;;) {
                bool fix = false;
                uint32_t want = 0;
                if (active && t > 0) { want = S.s_exit[t - 1]; fix = want != entry; }
                if (!__syncthreads_or(fix)) break;
                if (fix) {
                    entry = want;
                    uint32_t pos = nominal + entry;
                    rd_seek(r, stage_a, pos);
                    count = advance(A, slut_a, r, pos, end);
                    exitv = pos - end;
                    S.s_exit[t] = exitv;
                }
                __syncthreads();
            }
        }
        if (active) rec[P.sub_base[k] + i] = pack_rec(entry, exitv, count);
        uint32_t c = active ? count : 0;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
        if (lane == 0) S.s_red[wid] = c;
        __syncthreads();
        if (t == 0) {
            uint32_t a = 0;
            for (int w = 0; w < DT / 32; ++w) a += S.s_red[w];
            seqcnt[P.seq_base[k] + sq] = a;
        }
        carry_exit = S.s_exit[DT - 1];
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// fix kernel: CTA boundaries + scan of sequence counts.  Reads the stream from global memory.
// ---------------------------------------------------------------------------------------------
struct GRd {                  // bit reader over global memory, zero bits outside the chunk
    const uint8_t* base; uint32_t csize; uint32_t hi, lo, sh; int64_t next;
};
__device__ __forceinline__ uint32_t g_word(const GRd& r, int64_t boff) {
    uint32_t w = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int64_t b = boff + j;
        const uint32_t v = (b >= 0 && b < (int64_t)r.csize) ? r.base[b] : 0;
        w = (w << 8) | v;
    }
    return w;
}
__device__ __forceinline__ void g_seek(GRd& r, uint64_t bitpos) {
    const int64_t b = (int64_t)(bitpos >> 5) * 4;
    r.sh = (uint32_t)(bitpos & 31);
    r.hi = g_word(r, b); r.lo = g_word(r, b + 4); r.next = b + 8;
}
__device__ __forceinline__ void g_skip(GRd& r, uint32_t l) {
    r.sh += l;
    if (r.sh >= 32) { r.sh -= 32; r.hi = r.lo; r.lo = g_word(r, r.next); r.next += 4; }
}
// counts the codewords that begin in [start, nominal + SUB_BITS) (start >= nominal); exit offset
__device__ __forceinline__ void g_scan(const DecAux& A, const uint16_t* __restrict__ slut, GRd& r, uint64_t start,
                                       uint64_t nominal, uint32_t* count, uint32_t* exitv) {
    uint64_t pos = start;
    const uint64_t end = nominal + DEC_SUB_BITS;
    g_seek(r, pos);
    uint32_t cnt = 0;
    while (pos < end) {
        const uint32_t v = __funnelshift_l(r.lo, r.hi, r.sh);
        const uint32_t e = slut[v >> (32 - LUTB)];
        uint32_t l = (e >> 6) & 63;
        if ((e >> 12) == 0) { l = long_len(A, v, e & 63, l); if (!l) l = 1; }
        g_skip(r, l); pos += l; ++cnt;
    }
    *count = cnt; *exitv = (uint32_t)(pos - end);
}

__global__ void __launch_bounds__(DT)
dec_fix_kernel(const uint8_t* __restrict__ comp, const uint64_t* __restrict__ comp_off,
               const uint32_t* __restrict__ comp_size, DecPlan P, const uint8_t* __restrict__ tables,
               uint32_t* __restrict__ rec, uint32_t* __restrict__ seqcnt, int* status) {
    __shared__ __align__(16) uint16_t slut[LUTN];
    __shared__ __align__(16) uint8_t aux_raw[1024];
    __shared__ uint32_t s_warp[DT / 32 + 1];
    const uint32_t k = blockIdx.x, t = threadIdx.x, lane = t & 31, wid = t >> 5;
    const uint32_t nsub = P.nsub[k];
    if (nsub == 0) return;
    const uint32_t nseq = (nsub + DT - 1) / DT;
    const uint32_t ncta = (nseq + DEC_SEQ_PER_CTA - 1) / DEC_SEQ_PER_CTA;
    uint32_t* R = rec + P.sub_base[k];
    uint32_t* SC = seqcnt + P.seq_base[k];

    if (ncta > 1) {
        bool any = false;
        for (uint32_t b = 1 + t; b < ncta; b += DT) {
            const uint32_t i = b * DEC_SUBS_PER_CTA;
            any |= ((R[i - 1] >> 8) & 0xFF) != (R[i] & 0xFF);
        }
        if (__syncthreads_or(any)) {
            const uint8_t* tb = tables + (size_t)P.tab_idx[k] * DEC_TABLE_BYTES;
            copy_g2s16(slut, tb + DEC_TAB_S, LUTN * 2);
            copy_g2s16(aux_raw, tb + DEC_TAB_AUX, 1024);
            __syncthreads();
            const DecAux& A = *reinterpret_cast<const DecAux*>(aux_raw);
            if (A.bad) return;
            GRd r;
            r.base = comp + comp_off[k];
            r.csize = comp_size[k];
            for (;;) {
                bool changed = false;
                for (uint32_t b0 = 1; b0 < ncta; b0 += DT) {
                    const uint32_t b = b0 + t;
                    uint32_t want = 0, i = 0;
                    bool walk = false;
                    if (b < ncta) {
                        i = b * DEC_SUBS_PER_CTA;
                        want = (R[i - 1] >> 8) & 0xFF;
                        walk = want != (R[i] & 0xFF);
                    }
                    __syncthreads();           // all reads of neighbours' exits before any update
                    if (walk) {
                        const uint32_t iend = min(nsub, i + DEC_SUBS_PER_CTA);
                        uint32_t entry = want;
                        for (; i < iend; ++i) {
                            const uint64_t nominal = (uint64_t)i * DEC_SUB_BITS;
                            uint32_t count, exitv;
                            g_scan(A, slut, r, nominal + entry, nominal, &count, &exitv);
                            const uint32_t old = R[i];
                            R[i] = pack_rec(entry, exitv, count);
                            atomicAdd(&SC[i / DT], count - (old >> 16));
                            if (exitv == ((old >> 8) & 0xFF)) break;     // re-synchronised
                            entry = exitv;
                            if (i + 1 == iend) changed = true;            // ran off the CTA range
                        }
                    }
                    __syncthreads();
                }
                if (!__syncthreads_or(changed)) break;
            }
        }
    }
    __syncthreads();
    // exclusive scan of the sequence counts (in place)
    uint32_t carry = 0;
    for (uint32_t base = 0; base < nseq; base += DT) {
        const uint32_t i = base + t;
        uint32_t v = i < nseq ? SC[i] : 0, inc = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t o = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += o;
        }
        if (lane == 31) s_warp[wid] = inc;
        __syncthreads();
        if (t == 0) {
            uint32_t a = 0;
            for (int w = 0; w < DT / 32; ++w) { uint32_t x = s_warp[w]; s_warp[w] = a; a += x; }
            s_warp[DT / 32] = a;
        }
        __syncthreads();
        if (i < nseq) SC[i] = carry + s_warp[wid] + inc - v;
        carry += s_warp[DT / 32];
        __syncthreads();
    }
    (void)status;
}

// ---------------------------------------------------------------------------------------------
// write kernel
// ---------------------------------------------------------------------------------------------
// The write kernel's CTA is `groups` (1..3) independent 256-thread groups that share the chunk's
// lookup table (32 KiB) and take the CTA's sequences round-robin; each group has its own staging
// buffer, mbarrier, named barrier and per-warp output windows.  Sharing the table is what lets
// 24 warps be resident per SM.
struct WriteShared {
    __align__(16) uint2 wlut[LUTN];
    __align__(16) uint8_t aux[1024];
    uint32_t s_k;
};
static_assert(offsetof(WriteShared, aux) + offsetof(DecAux, sorted) == DEC_W_SORTED_REL, "wlut long-code entries address sorted[] relative to wlut");
struct WriteGroup {                      // followed by DT/32 output windows of win_bytes each
    __align__(16) uint8_t stage[DEC_STAGE_BYTES];
    __align__(8) uint64_t bar;
    uint32_t s_warp[DT / 32 + 1];
};
#define DEC_WRITE_SHARED ((sizeof(WriteShared) + 15) & ~(size_t)15)
#define DEC_WRITE_GROUP ((sizeof(WriteGroup) + 15) & ~(size_t)15)
#define DEC_WRITE_MAX_GROUPS 3

__device__ __forceinline__ void wgroup_sync(uint32_t grp) {
    if (grp == 0) asm volatile("bar.sync 1, %0;" ::"n"(DT) : "memory");
    else if (grp == 1) asm volatile("bar.sync 2, %0;" ::"n"(DT) : "memory");
    else asm volatile("bar.sync 3, %0;" ::"n"(DT) : "memory");
}

__global__ void __launch_bounds__(DT * DEC_WRITE_MAX_GROUPS)
dec_write_kernel(const uint8_t* __restrict__ comp, uint64_t comp_bytes, const uint64_t* __restrict__ comp_off,
                 const uint32_t* __restrict__ comp_size, const uint32_t* __restrict__ orig_size,
                 const uint8_t* __restrict__ len_tab, uint32_t K, DecPlan P, const uint8_t* __restrict__ tables,
                 const uint32_t* __restrict__ rec, const uint32_t* __restrict__ seqoff,
                 uint8_t* __restrict__ out, uint64_t out_cap, uint32_t win_bytes, uint32_t bulk_out, int* status) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    WriteShared& W = *reinterpret_cast<WriteShared*>(smem_raw);
    const uint32_t grp = threadIdx.x / DT, ngrp = blockDim.x / DT;
    const uint32_t t = threadIdx.x % DT, lane = t & 31, wid = t >> 5;             // thread index within the group
    const size_t gbytes = DEC_WRITE_GROUP + (DT / 32) * (size_t)win_bytes;
    uint8_t* gsm = smem_raw + DEC_WRITE_SHARED + grp * gbytes;
    WriteGroup& S = *reinterpret_cast<WriteGroup*>(gsm);
    if (blockIdx.x >= P.cta_base[K]) return;
    if (t == 0) {
        if (grp == 0) W.s_k = find_chunk(P.cta_base, K, blockIdx.x);
        mbar_init(&S.bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const uint32_t k = W.s_k;
    const uint32_t osize = orig_size[k];
    const uint64_t ooff = P.orig_off[k];
    if (ooff + osize > out_cap) { if (threadIdx.x == 0) hz_set_status(status, HZ_ERR_OUT_TOO_SMALL); return; }
    DecAux& A = *reinterpret_cast<DecAux*>(W.aux);
    const uint32_t ti = P.tab_idx[k];
    if (ti == DEC_NO_TABLE) {
        if (grp == 0) build_tables<true, false>(A, W.wlut, nullptr, S.stage, len_tab + (size_t)k * 256);
    } else {
        const uint8_t* tb = tables + (size_t)ti * DEC_TABLE_BYTES;
        copy_g2s16(W.wlut, tb + DEC_TAB_W, LUTN * 8);
        copy_g2s16(W.aux, tb + DEC_TAB_AUX, 1024);
    }
    __syncthreads();
    if (A.bad) { if (threadIdx.x == 0) hz_set_status(status, HZ_ERR_BAD_LENGTHS); return; }

    const uint64_t coff = comp_off[k];
    const uint32_t csize = comp_size[k];
    const uint32_t nsub = P.nsub[k];
    const uint32_t nseq = (nsub + DT - 1) / DT;
    const uint32_t cta_in_chunk = blockIdx.x - P.cta_base[k];
    const uint32_t sq0 = cta_in_chunk * DEC_SEQ_PER_CTA;
    const uint32_t nq = min((uint32_t)DEC_SEQ_PER_CTA, nseq - sq0);
    const uint64_t gout = reinterpret_cast<uint64_t>(out) + ooff;      // address of the chunk's first output byte
    uint8_t* win = gsm + DEC_WRITE_GROUP + (size_t)wid * win_bytes;
    const uint32_t win_a = pin_reg(smem_u32(win)), wlut_a = pin_reg(smem_u32(W.wlut)), aux_a = pin_reg(smem_u32(W.aux));
    const uint32_t stage_a = pin_reg(smem_u32(S.stage));
    uint32_t phase = 0;                                                // parity of the group's mbarrier
    const uint32_t sub0 = P.sub_base[k], seq0 = P.seq_base[k];
    uint32_t rv_next = 0, sbase_next = 0;
    if (grp < nq) {
        const uint32_t i1 = (sq0 + grp) * DT + t;
        rv_next = i1 < nsub ? rec[sub0 + i1] : 0;
        sbase_next = seqoff[seq0 + sq0 + grp];
    }

    // the compressed bytes of a group's NEXT sequence are requested as soon as every thread of the group
    // has finished reading the current ones (before the windows are copied out), so the bulk copy flies
    // during the copy-out and the next sequence's offset scan
    if (t == 0 && grp < nq) stage_issue(S.stage, &S.bar, stage_geom(comp, comp_bytes, coff, csize, sq0 + grp));
    for (uint32_t q = grp; q < nq; q += ngrp) {
        const uint32_t sq = sq0 + q;
        const StageGeom g = stage_geom(comp, comp_bytes, coff, csize, sq);
        bool released = false;                                         // this warp has passed the "stage is free" barrier
        // while the copy is in flight: output offsets of this sequence
        const uint32_t i = sq * DT + t;
        const bool active = i < nsub;
        const uint32_t rv = rv_next;
        const uint32_t sbase = sbase_next;
        if (q + ngrp < nq) {                                           // next sequence of this group: loads fly during this one
            const uint32_t i2 = (sq + ngrp) * DT + t;
            rv_next = i2 < nsub ? rec[sub0 + i2] : 0;
            sbase_next = seqoff[seq0 + sq + ngrp];
        }
        const uint32_t count = rv >> 16;
        uint32_t inc = count;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t x = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += x;
        }
        if (lane == 31) S.s_warp[wid] = inc;
        wgroup_sync(grp);
        // symbols of the warps before this one: every warp scans the DT/32 warp totals itself (one barrier)
        uint32_t wsum = lane < DT / 32 ? S.s_warp[lane] : 0;
#pragma unroll
        for (int d = 1; d < DT / 32; d <<= 1) {
            uint32_t x = __shfl_up_sync(0xffffffffu, wsum, d);
            if (lane >= d) wsum += x;
        }
        const uint32_t wbase = __shfl_sync(0xffffffffu, wsum, wid) - __shfl_sync(0xffffffffu, lane < DT / 32 ? S.s_warp[lane] : 0, wid);
        const uint32_t obase = sbase + wbase + inc - count;
        // the last subsequence of the chunk runs until orig_size symbols exist
        uint32_t todo = 0;
        if (active) {
            if (i == nsub - 1) {
                todo = osize > obase ? osize - obase : 0;
                if (todo > count) { hz_set_status(status, HZ_ERR_DECODE); todo = count; }   // fewer symbols than orig_size
            } else if (obase < osize) {
                todo = obase + count > osize ? osize - obase : count;
            }
        }
        mbar_wait(&S.bar, phase); phase ^= 1;
        stage_fixup(S.stage, g, t);
        wgroup_sync(grp);

        // ---- per-warp windowed decode -------------------------------------------------------
        // Packed counter C: bits 0-15 = stream bit position relative to a multiple of 32 at or before
        // the subsequence's first codeword (so C & 31 is the funnel-shift amount of the bit reader),
        // bits 16-30 = output bits produced in this pass (8 per symbol, starting at 8 * head).  One
        // add of the table entry's .y advances both; C2 > Clim catches long codes (bit 31) and
        // entries with more symbols than the pass may still take.
        BitRd r;
        const uint32_t start = g.bit0 + t * DEC_SUB_BITS + (rv & 0xFF);
        rd_seek(r, stage_a, start);
        uint32_t nx = lds32(r.wa + 4);                              // stage word after (hi, lo), NOT yet byte-swapped: loaded one refill ahead
        uint32_t rel = start & 31;
        uint64_t my_addr = gout + obase;                            // address of this lane's next symbol
        // the warp's output range: the lanes' runs are consecutive, so it is [first active lane's start, last active lane's end)
        uint64_t ws = ~0ull, we = 0ull;
        {
            const uint32_t am = __ballot_sync(0xffffffffu, todo != 0);
            if (am) {
                ws = __shfl_sync(0xffffffffu, my_addr, __ffs(am) - 1);
                we = __shfl_sync(0xffffffffu, my_addr + todo, 31 - __clz(am));
            }
        }
        for (uint64_t wa = ws & ~(uint64_t)15; wa < we; wa += win_bytes) {
            if (bulk_out) {                      // the previous bulk copy out of this window must have read it
                if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                __syncwarp();
            }
            uint32_t n = 0;
            if (todo && my_addr < wa + win_bytes) {
                const uint64_t room = wa + win_bytes - my_addr;
                n = todo < room ? todo : (uint32_t)room;
            }
            if (n) {
                const uint32_t woff = (uint32_t)(my_addr - wa);     // byte offset of the next symbol in the window
                const uint32_t w0 = win_a + (woff & ~3u);           // the word that holds it
                const uint32_t head = woff & 3;
                const uint32_t T8 = head + n;                       // output bytes of this pass, counted from the word's start
                uint32_t C = rel | (head << 19);
                const uint32_t Cend = T8 << 19;
                const uint32_t Cmain = T8 > 3 ? (T8 - 3) << 19 : 0;   // below it a whole entry (<= 4 symbols) always fits
                uint32_t sp = w0, acc = 0;
                // exactly ONE symbol at the reader's position (long codes, and the last <= 3 symbols of a pass)
                auto one_symbol = [&](uint32_t v, uint2 e, uint32_t& sym) -> uint32_t {
                    uint32_t l;
                    if (e.y >> 30 == 2) {                           // code longer than LUTB bits, one candidate length
                        sym = lds8(wlut_a + e.x + __funnelshift_l(v, 0u, e.y));   // + first l bits of v (l < 32)
                        return e.y & 0x7FFFFFFFu;
                    }
                    if (e.y >> 31) {                                // several candidate lengths, or no code
                        const uint32_t lmin = e.x & 63, lmax = (e.x >> 6) & 63;
                        l = lmin == lmax ? lmin : long_len(A, v, lmin, lmax);
                        if (l) {
                            const uint32_t sb = lds32(aux_a + (uint32_t)offsetof(DecAux, symbase) + l * 4);
                            sym = lds8(aux_a + (uint32_t)offsetof(DecAux, sorted) + sb + (v >> (32 - l)));
                        } else { l = 1; sym = 0; hz_set_status(status, HZ_ERR_DECODE); }
                    } else {
                        sym = e.x & 0xFF;
                        l = lds8(aux_a + (uint32_t)offsetof(DecAux, len) + sym);
                    }
                    return l + (8u << 16);
                };
                // one table lookup (up to four symbols): returns the advanced counter, refills the reader
                auto lookup = [&](uint32_t& syms) -> uint32_t {
                    const uint32_t v = rd_peek32(r, C);
                    const uint2 e = lds64(wlut_a + ((v >> 17) & 0x7FF8u));
                    syms = e.x;
                    uint32_t C2 = C + e.y;
                    if ((int32_t)C2 < 0) C2 = C + one_symbol(v, e, syms);
                    if ((C ^ C2) & 32) { r.hi = r.lo; r.lo = bswap32(nx); r.wa += 4; nx = lds32(r.wa + 4); }   // the word after next is already in flight
                    return C2;
                };
                if (head) {
                    // the first word is shared with the previous lane: its bytes are stored one by one
                    bool flushed = false;
                    while (C < Cmain && !flushed) {
                        uint32_t syms;
                        const uint32_t C2 = lookup(syms);
                        const uint32_t F = C >> 16;
                        acc |= __funnelshift_l(0u, syms, F);
                        if ((C ^ C2) & (32u << 16)) {
                            for (uint32_t j = head; j < 4; ++j) sts8(w0 + j, acc >> (8 * j));
                            acc = __funnelshift_l(syms, 0u, F);
                            sp = w0 + 4;
                            flushed = true;
                        }
                        C = C2;
                    }
                }
                if (C < Cmain) {
                    // Main loop (hand-written PTX, two lookups per trip so that the packed counter ping-pongs
                    // between two registers).  Per lookup: peek, table entry, advanced counter; a single-length
                    // long code (.y bit 31) is resolved by predicated instructions, no branch; then the predicated
                    // reader refill and output-word store.  ONE compare catches both the pass's last whole-entry
                    // lookup and the entries with several candidate lengths / no code (their bit 30 survives):
                    // those branch out of line, are resolved by the length search and re-enter the same
                    // iteration, so the warp reconverges every lookup.
#define HZ_WCOMMIT(CI, CO)                                                    \
    "xor.b32 t, " CI ", " CO ";\n"                                            \
    "and.b32 a, t, 32;\n"                                                     \
    "setp.ne.u32 p0, a, 0;\n"                                                 \
    "and.b32 a, t, 0x200000;\n"                                               \
    "setp.ne.u32 p1, a, 0;\n"                                                 \
    HZ_SHR16("F", CI)                                                         \
    "@p0 mov.u32 %1, %2;\n"                                                   \
    "@p0 prmt.b32 %2, %3, z, 0x0123;\n"                                       \
    "@p0 ld.shared.u32 %3, [%4+8];\n"                                         \
    HZ_ADD4("@p0", "%4", "%14")                                               \
    "shf.l.wrap.b32 a, z, ex, F;\n"                                           \
    "shf.l.wrap.b32 t, ex, z, F;\n"                                           \
    "add.u32 %5, %5, a;\n"                                                    \
    "@p1 st.shared.u32 [%6], %5;\n"                                           \
    HZ_ADD4("@p1", "%6", "%14")                                               \
    "selp.b32 %5, t, %5, p1;\n"
#define HZ_WSTEP(CI, CO, SFX)                                                 \
    "shf.l.wrap.b32 v, %2, %1, " CI ";\n"                                     \
    HZ_SHR20("ix", "v")                                                       \
    "mad.lo.u32 ix, ix, 8, %8;\n"                                             \
    "ld.shared.v2.u32 {ex, ey}, [ix];\n"                                      \
    "shf.l.wrap.b32 t, v, z, ey;\n"                                           \
    "add.u32 a, ex, %8;\n"                                                    \
    "add.u32 a, a, t;\n"                                                      \
    "add.u32 " CO ", " CI ", ey;\n"                                           \
    "setp.lt.s32 pl, ey, 0;\n"                                                \
    "@pl ld.shared.u8 ex, [a];\n"                                             \
    "@pl add.u32 " CO ", " CO ", 0x80000000;\n"                               \
    "setp.ge.u32 px, " CO ", %7;\n"                                           \
    "@px bra HZW_CHECK" SFX ";\n"                                             \
    "HZW_BACK" SFX ":\n"                                                      \
    HZ_WCOMMIT(CI, CO)
#define HZ_WCHECK(CI, CO, SFX, FIN)                                           \
    "HZW_CHECK" SFX ":\n"                                                     \
    "and.b32 t, " CO ", 0x40000000;\n"                                        \
    "setp.eq.u32 pq, t, 0;\n"                                                 \
    "@pq bra HZW_LAST" SFX ";\n"                                              \
    "ld.shared.v2.u32 {ex, ey}, [ix];\n"                                      \
    "and.b32 l, ex, 63;\n"                                                    \
    "shr.u32 m, ex, 6;\n"                                                     \
    "and.b32 m, m, 63;\n"                                                     \
    "add.u32 auxb, %8, %10;\n"                                                \
    HZ_PTX_LONGLEN("W" SFX)                                                   \
    "setp.eq.u32 pq, l, 0;\n"                                                 \
    "@pq bra HZW_BAD" SFX ";\n"                                               \
    "mad.lo.u32 a, l, 4, auxb;\n"                                             \
    "ld.shared.u32 t, [a+%11];\n"                                             \
    "sub.u32 u, 32, l;\n"                                                     \
    "shr.u32 u, v, u;\n"                                                      \
    "add.u32 a, t, u;\n"                                                      \
    "add.u32 a, a, auxb;\n"                                                   \
    "ld.shared.u8 ex, [a+%12];\n"                                             \
    "bra HZW_GOT" SFX ";\n"                                                   \
    "HZW_BAD" SFX ":\n"                                                       \
    "mov.u32 l, 1;\n"                                                         \
    "mov.u32 ex, 0;\n"                                                        \
    "atom.global.cas.b32 t, [%9], 0, %13;\n"                                  \
    "HZW_GOT" SFX ":\n"                                                       \
    "add.u32 " CO ", " CI ", l;\n"                                            \
    "add.u32 " CO ", " CO ", 0x80000;\n"                                      \
    "setp.lt.u32 pq, " CO ", %7;\n"                                           \
    "@pq bra HZW_BACK" SFX ";\n"                                              \
    "HZW_LAST" SFX ":\n"                                                      \
    HZ_WCOMMIT(CI, CO)                                                        \
    FIN                                                                       \
    "bra HZW_DONE;\n"
                    asm volatile(
                        "{\n"
                        ".reg .pred pl, px, p0, p1, pq;\n"
                        ".reg .u32 D, v, ix, ex, ey, t, u, a, F, z, l, m, auxb;\n"
                        "mov.u32 z, 0;\n"
                        "HZW_TOP:\n"
                        HZ_WSTEP("%0", "D", "1")
                        HZ_WSTEP("D", "%0", "2")
                        "bra HZW_TOP;\n"
                        HZ_WCHECK("%0", "D", "1", "mov.u32 %0, D;\n")
                        HZ_WCHECK("D", "%0", "2", "")
                        "HZW_DONE:\n"
                        "}\n"
                        : "+r"(C), "+r"(r.hi), "+r"(r.lo), "+r"(nx), "+r"(r.wa), "+r"(acc), "+r"(sp)
                        : "r"(Cmain), "r"(smem_u32(W.wlut)), "l"(status),
                          "n"(LUTN * 8), "n"(offsetof(DecAux, symbase)), "n"(offsetof(DecAux, sorted)), "n"(HZ_ERR_DECODE)
                          HZ_W_TWO_OPERAND
                        : "memory");
#undef HZ_WSTEP
#undef HZ_WCHECK
#undef HZ_WCOMMIT
                }
                // bytes still in the accumulator, then the pass's last symbols one at a time
                const uint32_t pend = (C >> 19) & 3;
                for (uint32_t j = sp == w0 ? head : 0; j < pend; ++j) sts8(sp + j, acc >> (8 * j));
                uint32_t bp = sp + pend;
                while (C < Cend) {
                    const uint32_t v = rd_peek32(r, C);
                    const uint2 e = lds64(wlut_a + ((v >> 17) & 0x7FF8u));
                    uint32_t sym;
                    const uint32_t C2 = C + one_symbol(v, e, sym);
                    if ((C ^ C2) & 32) { r.hi = r.lo; r.lo = bswap32(nx); r.wa += 4; nx = lds32(r.wa + 4); }   // the word after next is already in flight
                    sts8(bp, sym); ++bp;
                    C = C2;
                }
                rel = C & 0xFFFFu;
                todo -= n; my_addr += n;
            }
            __syncwarp();
            if (wa + win_bytes >= we) {          // this warp's last pass: it no longer reads the staged bytes
                wgroup_sync(grp);
                released = true;
                if (t == 0 && q + ngrp < nq) stage_issue(S.stage, &S.bar, stage_geom(comp, comp_bytes, coff, csize, sq + ngrp));
            }
            // copy the window out: aligned 16-byte units; the ragged first / last unit byte by byte,
            // one byte per lane (lanes 0-15: first unit, lanes 16-31: last unit)
            const uint64_t lo = ws > wa ? ws : wa;
            const uint64_t hi = we < wa + win_bytes ? we : wa + win_bytes;
            const uint32_t b0 = (uint32_t)(lo - wa), b1 = (uint32_t)(hi - wa);          // byte range [b0, b1) of the window
            const uint32_t f0 = (b0 + 15) >> 4, f1 = b1 >> 4;                              // whole units [f0, f1)
            if (bulk_out) {
                // whole units leave with ONE asynchronous bulk copy (shared -> global) issued by lane 0: the warp
                // goes on to its next pass / sequence while the copy drains
                if (f1 > f0) {
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) {
                        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                                     ::"l"(wa + (uint64_t)f0 * 16), "r"(win_a + f0 * 16), "r"((f1 - f0) * 16) : "memory");
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                }
            } else {
                for (uint32_t u = f0 + lane; u < f1; u += 32)
                    *reinterpret_cast<uint4*>(wa + (uint64_t)u * 16) = *reinterpret_cast<const uint4*>(win + u * 16);
            }
            {
                const uint32_t bb = lane < 16 ? (b0 & ~15u) + lane : (b1 & ~15u) + (lane - 16);
                const bool edge = lane < 16 ? (b0 & 15) != 0 : (b1 & 15) != 0;
                if (edge && bb >= b0 && bb < b1) *reinterpret_cast<uint8_t*>(wa + bb) = win[bb];
            }
            __syncwarp();
        }
        if (!released) {                         // a warp without output still takes part in the hand-over
            wgroup_sync(grp);
            if (t == 0 && q + ngrp < nq) stage_issue(S.stage, &S.bar, stage_geom(comp, comp_bytes, coff, csize, sq + ngrp));
        }
    }
    if (bulk_out && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // shared memory must outlive the copies
}

// ---------------------------------------------------------------------------------------------
int hzk_decode(hz_ctx* ctx, const uint8_t* d_comp, uint64_t comp_bytes, const uint64_t* d_comp_off,
               const uint32_t* d_comp_size, const uint32_t* d_orig_size, const uint64_t* d_orig_off,
               const uint8_t* d_len, uint32_t K, uint8_t* d_out, uint64_t out_cap) {
    if (K == 0) return HZ_OK;
    // plan arrays
    HZ_TRY(hz_reserve(ctx, &ctx->dec_meta, ((size_t)K + 1) * (6 * sizeof(uint32_t) + sizeof(uint64_t)) + 64));
    DecPlan P;
    uint8_t* m = (uint8_t*)ctx->dec_meta.p;
    P.orig_off = (uint64_t*)m; m += ((size_t)K + 1) * sizeof(uint64_t);
    P.nsub = (uint32_t*)m; m += ((size_t)K + 1) * sizeof(uint32_t);
    P.sub_base = (uint32_t*)m; m += ((size_t)K + 1) * sizeof(uint32_t);
    P.seq_base = (uint32_t*)m; m += ((size_t)K + 1) * sizeof(uint32_t);
    P.cta_base = (uint32_t*)m; m += ((size_t)K + 1) * sizeof(uint32_t);
    P.tab_idx = (uint32_t*)m; m += ((size_t)K + 1) * sizeof(uint32_t);
    P.islice = (uint32_t*)m;
    // Chunks decoded by a single CTA used to build their lookup tables inside the sync AND the write kernel
    // (two builds per chunk, each on a CTA that holds the decode kernels' shared memory): with thousands of
    // small chunks that was 60 % of the decode time.  While K tables fit DEC_TAB_PREBUILD_CAP bytes of scratch
    // every chunk's tables are built once by dec_tables_kernel (a small CTA, many per SM) and copied in.
    uint32_t tab_min_seq = DEC_SEQ_PER_CTA;
    if ((uint64_t)K * DEC_TABLE_BYTES <= DEC_TAB_PREBUILD_CAP) tab_min_seq = 0;
    if (ctx->knobs.dec_prebuild >= 0) tab_min_seq = ctx->knobs.dec_prebuild ? 0 : DEC_SEQ_PER_CTA;   // developer knob
    HZ_TRY(hz_reserve(ctx, &ctx->dec_misc, (size_t)K + 64));
    uint8_t* ident = (uint8_t*)ctx->dec_misc.p;
    const uint32_t ident_on = ctx->knobs.ident;   // developer knob: HZ_IDENT=0 sends identity chunks through the table walk
    HZ_LAUNCH(ctx, "dec_ident", dec_ident_flags_kernel, (K + DT / 32 - 1) / (DT / 32), DT, 0, d_len, d_orig_size, K, ident, ident_on);
    // Every stream goes to the fused single-walk kernel (hz_decode_fused.cu), whose CTA shape follows the units per
    // chunk; the multi-pass kernels below remain as the developer knob HZ_DEC=legacy (A/B and a second implementation
    // for the parity tests).  Measured on B200 at 4 bits/symbol, fused vs multi-pass, GB/s of output: 1 KiB chunks
    // 23.5 vs 7.2, 4 KiB 89 vs 29, 16 KiB 308 vs 130, 64 KiB 579 vs 384, 256 KiB 762 vs 572, 1 MiB 958 vs 647, 16 MiB 886 vs 711.
    const int mode = ctx->knobs.dec_mode;   // developer knob HZ_DEC=legacy|fused
    const bool fused = mode != 1;
    if (fused) {
        const uint64_t* p_orig_off = nullptr; const uint32_t* p_islice = nullptr;
        HZ_TRY(hzk_decode_fused(ctx, d_comp, comp_bytes, d_comp_off, d_comp_size, d_orig_size, d_orig_off, d_len, K, d_out, out_cap,
                                ident, &p_orig_off, &p_islice));
        HZ_LAUNCH(ctx, "dec_ident_copy", dec_ident_copy_kernel, 8 * ctx->sm_count, DT, 0, d_comp, d_comp_off, d_comp_size,
                  d_orig_size, p_orig_off, p_islice, K, d_out, out_cap, ctx->d_status);
        return HZ_OK;
    }
    HZ_LAUNCH(ctx, "dec_plan", dec_plan_kernel, 1, 1024, 0, d_comp_off, comp_bytes, d_comp_size, d_orig_size, d_orig_off, K, P,
              tab_min_seq, ident, ctx->d_status);
    // upper bounds (no host sync): every chunk has at most ceil(comp_size*8/SUB_BITS)+1 subsequences
    const uint64_t max_sub = comp_bytes * 8 / DEC_SUB_BITS + 2ull * K + 2;
    const uint64_t max_seq = max_sub / DT + K + 1;
    const uint64_t max_cta = max_seq / DEC_SEQ_PER_CTA + K + 1;
    uint64_t max_tab = comp_bytes / ((uint64_t)DEC_SEQ_PER_CTA * DEC_SEQ_BYTES) + 1;
    if (max_tab > K || tab_min_seq == 0) max_tab = K;
    if (max_cta > 0x7fffffffull) return hz_fail(ctx, HZ_ERR_ARG, "decode grid too large");
    HZ_TRY(hz_reserve(ctx, &ctx->dec_rec, max_sub * sizeof(uint32_t)));
    HZ_TRY(hz_reserve(ctx, &ctx->dec_seqcnt, max_seq * sizeof(uint32_t)));
    HZ_TRY(hz_reserve(ctx, &ctx->dec_tables, max_tab * DEC_TABLE_BYTES));
    uint32_t* rec = (uint32_t*)ctx->dec_rec.p;
    uint32_t* seqcnt = (uint32_t*)ctx->dec_seqcnt.p;
    uint8_t* tables = (uint8_t*)ctx->dec_tables.p;
    if (!ctx->attr_decode) {
        HZ_CUDA(ctx, cudaFuncSetAttribute(dec_sync_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SyncSmem)));
        HZ_CUDA(ctx, cudaFuncSetAttribute(dec_write_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 1024));
        ctx->attr_decode = true;
    }
    HZ_LAUNCH(ctx, "dec_ident_copy", dec_ident_copy_kernel, 8 * ctx->sm_count, DT, 0, d_comp, d_comp_off, d_comp_size,
              d_orig_size, P.orig_off, P.islice, K, d_out, out_cap, ctx->d_status);
    HZ_LAUNCH(ctx, "dec_tables", dec_tables_kernel, K, DT, 0, d_len, P, tables, ctx->d_status);
    HZ_LAUNCH(ctx, "dec_sync", dec_sync_kernel, (unsigned)max_cta, DT, sizeof(SyncSmem),
              d_comp, comp_bytes, d_comp_off, d_comp_size, d_len, K, P, tables, rec, seqcnt, ctx->d_status);
    HZ_LAUNCH(ctx, "dec_fix", dec_fix_kernel, K, DT, 0,
              d_comp, d_comp_off, d_comp_size, P, tables, rec, seqcnt, ctx->d_status);
    // per-warp output window: sized so that the symbols of one warp (32 subsequences) normally fit one
    // window, from the stream's overall expansion ratio; smaller windows -> more CTAs per SM
    uint64_t est = comp_bytes ? (uint64_t)(32.0 * DEC_SUB_BYTES * 1.1 * (double)out_cap / (double)comp_bytes) + 64 : DEC_WIN_MIN;
    est = (est + 255) & ~(uint64_t)255;
    uint32_t win_bytes = (uint32_t)(est < DEC_WIN_MIN ? DEC_WIN_MIN : (est > DEC_WIN_MAX ? DEC_WIN_MAX : est));
    if (ctx->knobs.dec_win >= 256 && ctx->knobs.dec_win <= DEC_WIN_MAX) win_bytes = (uint32_t)ctx->knobs.dec_win & ~255u;   // developer knob
    // groups per CTA: as many as fit the SM's shared memory next to the shared table (developer knob HZ_DEC_GROUPS)
    const size_t gbytes = DEC_WRITE_GROUP + (DT / 32) * (size_t)win_bytes;
    uint32_t groups = (uint32_t)((227 * 1024 - 1024 - DEC_WRITE_SHARED) / gbytes);
    if (groups > DEC_WRITE_MAX_GROUPS) groups = DEC_WRITE_MAX_GROUPS;
    if (groups < 1) groups = 1;
    // chunks of a few sequences each build their table inside the CTA: two independent CTAs per SM
    // (two table builds in flight) beat one CTA whose extra groups wait for the build
    const uint32_t fit = groups;
    if (comp_bytes / K < 4ull * DEC_SEQ_BYTES) groups = 1;
    if (ctx->knobs.dec_groups >= 1 && ctx->knobs.dec_groups <= (int)fit) groups = (uint32_t)ctx->knobs.dec_groups;
    const uint32_t bulk_out = ctx->knobs.dec_bulk;   // developer knob: HZ_DEC_BULK=0 copies the windows out with 128-bit stores
    HZ_LAUNCH(ctx, "dec_write", dec_write_kernel, (unsigned)max_cta, DT * groups, DEC_WRITE_SHARED + groups * gbytes,
              d_comp, comp_bytes, d_comp_off, d_comp_size, d_orig_size, d_len, K, P, tables, rec, seqcnt, d_out, out_cap,
              win_bytes, bulk_out, ctx->d_status);
    return HZ_OK;
}
