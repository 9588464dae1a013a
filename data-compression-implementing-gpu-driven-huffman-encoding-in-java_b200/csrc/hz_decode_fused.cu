// hz_decode_fused.cu — single-residency chunked Huffman decode: ONE walk over the compressed stream.
//
// Replaces CanonicalHuffman.generateCanonicalCodesFromLengths + TableBasedHuffmanDecoder.decode
// (core/CanonicalHuffman.java:141-146, core/TableBasedHuffmanDecoder.java:36-152, driven by
// CpuCompressionService.decodeChunkParallel, service/cpu/CpuCompressionService.java:511-532).
//
// A chunk of the .dcz payload is one sequential bitstream without restart markers.  It is cut into
// subsequences of S 32-bit words (S per chunk, so that a subsequence holds ~128 symbols); 32 consecutive
// subsequences form a UNIT, the work item of one warp.  A warp
//   1. receives its unit's bytes by a 1-D TMA bulk copy into its private stage (cp.async.bulk + mbarrier),
//      zero-fills what lies outside the chunk (units at a chunk's ends only; the reader byte-swaps words as it refills);
//   2. lane i starts a few words before subsequence i (a guess), walks to the first codeword boundary inside it
//      (self-synchronisation), then decodes the subsequence ONCE, up to four symbols per table lookup, into its
//      private row of shared memory (word-interleaved rows: lane == bank, no conflicts), counting symbols;
//      exit[i-1] == entry[i] is checked with shuffles and mismatching lanes re-walk from the neighbour's exit;
//   3. publishes {entry of lane 0, exit of lane 31, symbol count} and obtains the unit's output offset by a
//      decoupled look-back over the records of the chunk's earlier units.  A record is FINAL once its chain of
//      entry == previous exit links reaches the chunk's first unit; a unit whose own link is broken re-walks
//      from the true entry before it finalises, so the result never depends on the guesses;
//   4. compacts the 32 rows in place (all rows to registers, then shifted word stores) into one contiguous
//      window aligned like the destination and sends it to global memory with one asynchronous bulk copy.
// Units are handed out in stream order by a per-chunk atomic counter, so a unit only ever waits for units that
// running warps hold.  A CTA keeps one chunk's lookup table (33 KiB, delivered by one bulk copy, shared by its warps);
// CTAs take chunks from a global ticket (drawn a chunk ahead) and, when none are left, join chunks that still have
// units.  The CTA shape follows the stream's units per chunk - 24 warps x 1 CTA per SM, 8 x 2 or 5 x 3 - so that
// streams of small chunks keep several independent CTAs per SM (hzk_decode_fused); the chunk's last subsequence ends
// with the chunk, symbols the stream does not hold are the all-zero codeword's (TableBasedHuffmanDecoder.java:204-208).
#include <cstdio>
#include "hz_decode_tables.cuh"

#define FU_WARPS_MAX 24                            // large chunks: ONE CTA of 24 warps per SM shares a chunk's table
#define FU_SUB_MIN 3
#define FU_SUB_MAX 17
#define FU_ROW_WORDS 46                            // 44 words of symbols + 2 guard words (overflow is tested once per two lookups)
#define FU_CAP_SYMS 176
#define FU_CAP_BITS (FU_CAP_SYMS * 8)              // multiple of 32: the packed counter's low five bits stay the byte lane
#define FU_OUT_BIAS (0x8000u - FU_CAP_BITS)        // bit 15 of the output field <=> the row is full
#define FU_LEAD_MAX 8
#define FU_STAGE_BYTES 2256                        // 15 alignment + 32 lead-in + 17 * 128 + 32 look-ahead, rounded to 16
#define FU_ROWS_BYTES (32 * FU_ROW_WORDS * 4)
#define FU_WARP_BYTES (FU_STAGE_BYTES + FU_ROWS_BYTES)
#define FU_TABLE_BYTES (LUTN * 8 + 1024)           // wlut + DecAux
#define FU_NONE 0xFFFFFFFFu
#ifndef FU_NS0
#define FU_NS0 32                                  // look-back polls: first sleep and its cap, ns
#define FU_NSMAX 256
#endif
#define FU_RING 64                                 // > look-back window (32) + units in flight in one CTA (24)
#define FU_RING_CL 128                             // cluster pair: both CTAs' units of the shared chunk

template <int RN>
struct FuSharedT {
    __align__(16) uint2 wlut[LUTN];
    __align__(16) uint8_t aux[1024];
    __align__(16) uint4 ring[RN + (RN == FU_RING_CL ? 1 : 0)];   // look-back records: {record, unit + 1, chunk + 1} of the units THIS CTA decoded (and, in a
                                             // cluster, its peer); the cluster layout's extra entry holds peer_k (fu_peer_k): a lone CTA's layout is unchanged
    __align__(8) uint64_t bar[FU_WARPS_MAX];
    __align__(8) uint64_t tbar;              // the chunk's table arrives by ONE bulk copy
    uint32_t s_k, s_pick;
};
// cluster: chunk + 1 the PEER CTA works on (written by the peer through DSMEM), 0 = nothing announced yet
template <class SH> __device__ __forceinline__ volatile uint32_t* fu_peer_k(SH& S) { return reinterpret_cast<volatile uint32_t*>(&S.ring[FU_RING_CL]); }
typedef FuSharedT<FU_RING> FuShared;
static_assert(offsetof(FuShared, aux) + offsetof(DecAux, sorted) == DEC_W_SORTED_REL, "long-code entries address sorted[] relative to wlut");
static_assert(offsetof(FuShared, wlut) == 0 && offsetof(FuShared, aux) == LUTN * 8, "wlut + aux are one contiguous bulk-copy destination");
static_assert(offsetof(FuSharedT<FU_RING_CL>, aux) == LUTN * 8, "same table layout with the larger ring");
#define FU_SHARED_BYTES_RN(RN) ((sizeof(FuSharedT<RN>) + 15) & ~(size_t)15)
#define FU_SMEM_BYTES_CL(W, CL) (FU_SHARED_BYTES_RN((CL) == 2 ? FU_RING_CL : FU_RING) + (size_t)(W) * FU_WARP_BYTES)
#define FU_SMEM_BYTES(W) FU_SMEM_BYTES_CL(W, 1)
static_assert(FU_SMEM_BYTES_CL(24, 2) <= 227 * 1024, "24 warps + the cluster ring fit one SM");

struct FuPlan {
    uint64_t* orig_off;     // [K+1]
    uint32_t* nsub;         // [K]   subsequences (0: empty, identity or rejected chunk)
    uint32_t* nunit;        // [K]
    uint32_t* unit_base;    // [K+1] first look-back record of chunk k
    uint32_t* geom;         // [K]   S | lead-in words << 8
    uint32_t* unit_ctr;     // [K]   next unit of chunk k (device tickets)
    uint32_t* islice;       // [K+1] identity-copy slices before chunk k
    uint32_t* ctl;          // [0] chunk ticket
};
#define FU_IDENT_SLICE 65536u

// ---------------------------------------------------------------------------------------------
// plan: per-chunk geometry and prefix sums (1 CTA)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void fu_chunk_geom(uint32_t csize, uint32_t osize, bool ident, bool ok, uint32_t lead_knob,
                                              uint32_t& S, uint32_t& lead, uint32_t& ns) {
    S = FU_SUB_MAX; lead = 3; ns = 0;
    if (!ok || ident || osize == 0) return;
    // ~136 symbols per subsequence at most (a row holds 176): S <= 136 * (8 csize / osize) / 32 words, and ODD, so that
    // the 32 lanes' stream words (stride S) fall into 32 different shared-memory banks
    uint64_t s = ((uint64_t)csize * 34) / osize;
    s = s < FU_SUB_MIN ? FU_SUB_MIN : (s > FU_SUB_MAX ? FU_SUB_MAX : s);
    S = (uint32_t)((s - 1) | 1);
    // lead-in before a subsequence for the self-synchronisation guess: the nearer a code is to equal lengths, the
    // more codewords it takes to fall into step (measured with tests/fused_model.py on Zipf streams: a wrong guess
    // per 10^3..10^4 subsequences with 1 word at <= 2 bits/symbol, 2 words at 3..4, 6 at 6, 8 at 7)
    // x = 1.5 * bits per symbol (x4 fixed point): one word below 2.5 bits/symbol, two below 4.5, then 1.5 b - 3
    const uint64_t x4 = ((uint64_t)csize * 48) / osize;
    // (re-measured after the no-output walk went from 19 to 16 instructions per lookup: one word more from 4.5 bits/symbol
    //  on - 5 bits/symbol: 758 -> 784 GB/s with 6 words instead of 5, 6 bits/symbol: 629 -> 641 with 8 instead of 6)
    const int64_t l = lead_knob ? (int64_t)lead_knob
                                : (x4 < 15 ? 1 : (x4 < 21 ? 2 : (int64_t)((x4 + 2) / 4) - 3 + (x4 >= 27 ? 1 : 0)));
    lead = (uint32_t)(l < 1 ? 1 : (l > FU_LEAD_MAX ? FU_LEAD_MAX : l));
    if (lead > S) lead = S;
    const uint64_t bits = (uint64_t)csize * 8, sb = (uint64_t)S * 32;
    ns = (uint32_t)((bits + sb - 1) / sb);
    if (ns == 0) ns = 1;
}

// Rounds of 1024 consecutive chunks (coalesced loads, one chunk per thread), a block-wide exclusive scan of
// {units | identity slices << 32, original bytes} per round, running carries.
__device__ __forceinline__ uint64_t shfl_up64(uint64_t v, int d) {
    const uint32_t lo = __shfl_up_sync(0xffffffffu, (uint32_t)v, d), hi = __shfl_up_sync(0xffffffffu, (uint32_t)(v >> 32), d);
    return (uint64_t)lo | ((uint64_t)hi << 32);
}
__global__ void __launch_bounds__(1024)
fu_plan_kernel(const uint64_t* __restrict__ comp_off, const uint32_t* __restrict__ comp_size,
               const uint32_t* __restrict__ orig_size, const uint64_t* __restrict__ orig_off_in, uint64_t comp_bytes,
               uint32_t K, FuPlan P, const uint8_t* __restrict__ ident, uint32_t lead_knob, int* status) {
    __shared__ uint64_t wtot[2][2][32];                   // [round parity][value][warp]
    const uint32_t t = threadIdx.x, lane = t & 31, wid = t >> 5;
    uint64_t carry_a = 0, carry_b = 0;                    // units | slices << 32, original bytes: before this round
    uint32_t par = 0;
    for (uint32_t base = 0; base < K; base += 1024, par ^= 1) {
        const uint32_t i = base + t;
        uint64_t va = 0, vb = 0;
        uint32_t S = FU_SUB_MAX, lead = 3, ns = 0;
        if (i < K) {
            // the chunk must lie inside the addressable stream (untrusted footer fields reach this ABI)
            const uint64_t co = comp_off[i];
            const uint32_t cs = comp_size[i], os = orig_size[i];
            const bool id = ident[i] != 0;
            const bool ok = co <= comp_bytes && cs <= comp_bytes - co;
            if (!ok) hz_set_status(status, HZ_ERR_ARG);
            fu_chunk_geom(cs, os, id, ok, lead_knob, S, lead, ns);
            va = (uint64_t)((ns + 31) / 32) | ((uint64_t)((ok && id) ? (os + FU_IDENT_SLICE - 1) / FU_IDENT_SLICE : 0u) << 32);
            vb = os;
        }
        uint64_t ia = va, ib = vb;                        // inclusive scans within the warp
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint64_t xa = shfl_up64(ia, d), xb = shfl_up64(ib, d);
            if (lane >= (uint32_t)d) { ia += xa; ib += xb; }
        }
        if (lane == 31) { wtot[par][0][wid] = ia; wtot[par][1][wid] = ib; }
        __syncthreads();
        uint64_t pa = carry_a, pb = carry_b, ta = 0, tb = 0;
        for (uint32_t w = 0; w < 32; ++w) {
            const uint64_t xa = wtot[par][0][w], xb = wtot[par][1][w];
            if (w < wid) { pa += xa; pb += xb; }
            ta += xa; tb += xb;
        }
        if (i < K) {
            const uint64_t ea = pa + ia - va, eb = pb + ib - vb;   // exclusive prefixes of chunk i
            P.nsub[i] = ns; P.nunit[i] = (ns + 31) / 32; P.unit_base[i] = (uint32_t)ea; P.geom[i] = S | (lead << 8);
            P.unit_ctr[i] = 0;
            P.orig_off[i] = orig_off_in ? orig_off_in[i] : eb;
            P.islice[i] = (uint32_t)(ea >> 32);
        }
        carry_a += ta; carry_b += tb;
    }
    if (t == 0) {
        P.unit_base[K] = (uint32_t)carry_a; P.orig_off[K] = carry_b; P.islice[K] = (uint32_t)(carry_a >> 32);
        P.ctl[0] = 0;
    }
}

// look-back records of the units that exist (their number is only known on the device)
__global__ void __launch_bounds__(256)
fu_zero_kernel(uint64_t* __restrict__ rec, const uint32_t* __restrict__ unit_base, uint32_t K) {
    const uint32_t n = unit_base[K];
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) rec[i] = 0;
}

// tables: one CTA per chunk, written to global memory once, copied by every CTA that works on the chunk
__global__ void __launch_bounds__(DT, 6)
fu_tables_kernel(const uint8_t* __restrict__ len_tab, FuPlan P, uint8_t* __restrict__ tables) {
    __shared__ __align__(16) uint8_t scratch[DEC_BUILD_SCRATCH];
    __shared__ __align__(16) uint8_t aux_raw[1024];
    const uint32_t k = blockIdx.x;
    if (P.nsub[k] == 0) return;
    DecAux& A = *reinterpret_cast<DecAux*>(aux_raw);
    uint8_t* dst = tables + (size_t)k * FU_TABLE_BYTES;
    build_tables<true, false, 1>(A, reinterpret_cast<uint2*>(dst), nullptr, scratch, len_tab + (size_t)k * 256);
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < 1024 / 16; i += DT) reinterpret_cast<uint4*>(dst + LUTN * 8)[i] = reinterpret_cast<uint4*>(aux_raw)[i];
}

// ---------------------------------------------------------------------------------------------
// look-back records: state (2) | entry of lane 0 (6) | exit of lane 31 (6) | symbols (18) | inclusive prefix (32)
// ---------------------------------------------------------------------------------------------
#define FU_EMPTY 0u
#define FU_SPEC 1u
#define FU_FINAL 2u
__device__ __forceinline__ uint64_t fu_pack(uint32_t st, uint32_t entry, uint32_t exitv, uint32_t cnt, uint32_t pre) {
    return (uint64_t)(st | (entry << 2) | (exitv << 8) | (cnt << 14)) | ((uint64_t)pre << 32);
}
__device__ __forceinline__ uint64_t ld_rec(const uint64_t* p) {
    uint64_t v; asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ void st_rec(uint64_t* p, uint64_t v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ uint32_t warp_sum(uint32_t v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}
__device__ __forceinline__ uint32_t warp_excl_scan(uint32_t v, uint32_t lane) {
    uint32_t inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t x = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += x;
    }
    return inc - v;
}

__device__ __forceinline__ uint4 lds128(uint32_t a) {
    uint4 v; asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a)); return v;
}
__device__ __forceinline__ void sts128(uint32_t a, uint4 v) {
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// A record lives in global memory (every CTA that works on the chunk sees it) and, for the units this CTA decoded
// itself, in a shared-memory ring that spares the look-back the global round trips (a CTA that has a chunk to
// itself never leaves shared memory).  Ring entries are written and read as ONE 128-bit access and carry the unit and
// the chunk they belong to.  In a cluster pair (streams with fewer chunks than SMs, where CTAs share chunks from the
// start) a CTA also PUSHES its entries into the peer's ring through distributed shared memory while the peer works on
// the same chunk: the peer's look-backs then stay in its own shared memory instead of polling L2.
template <int CL> struct FuRingT;
template <> struct FuRingT<1> {      // a lone CTA: one register, the ring's shared address (chunk tag and peer are constants)
    uint32_t a;
};
template <> struct FuRingT<2> {
    uint32_t a;          // shared address of this CTA's ring
    uint32_t ktag;       // chunk + 1
    uint32_t peer;       // shared::cluster address of the peer's ring while it works on this chunk, else 0
};
template <int CL> __device__ __forceinline__ uint32_t fu_ktag(const FuRingT<CL>& g) { if constexpr (CL == 2) return g.ktag; else return 0u; }
__device__ __forceinline__ void fu_remote_st128(uint32_t ra, uint4 v) {
    asm volatile("st.shared::cluster.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(ra), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// (CL is the kernel's cluster size: for a lone CTA the ring size, the chunk tag and the peer are compile-time constants)
template <int CL>
__device__ __forceinline__ void fu_ring_store(const FuRingT<CL>& g, uint32_t u, uint4 e) {
    const uint32_t o = (u & (CL == 2 ? FU_RING_CL - 1 : FU_RING - 1)) * 16;
    sts128(g.a + o, e);
    if constexpr (CL == 2) { if (g.peer) fu_remote_st128(g.peer + o, e); }
}
template <int CL>
__device__ __forceinline__ void fu_put(uint64_t* R, const FuRingT<CL>& g, uint32_t u, uint64_t rec) {
    fu_ring_store<CL>(g, u, make_uint4((uint32_t)rec, (uint32_t)(rec >> 32), u + 1, fu_ktag<CL>(g)));
    st_rec(R + u, rec);
}
// A warp marks the unit it has just been handed as PENDING in the ring (tag set, record EMPTY): a look-back that meets
// the mark knows the record will appear HERE and polls shared memory - without it every look at a unit this CTA is
// still decoding went to global memory (a round trip to L2 to read EMPTY, and again for every poll).
template <int CL>
__device__ __forceinline__ void fu_mark_pending(const FuRingT<CL>& g, uint32_t u) {
    fu_ring_store<CL>(g, u, make_uint4(0u, 0u, u + 1, fu_ktag<CL>(g)));
}
template <int CL>
__device__ __forceinline__ uint64_t fu_get(const uint64_t* R, const FuRingT<CL>& g, int q) {
    const uint4 e = lds128(g.a + ((uint32_t)q & (CL == 2 ? FU_RING_CL - 1 : FU_RING - 1)) * 16);
    if (e.z == (uint32_t)q + 1 && (CL != 2 || e.w == fu_ktag<CL>(g))) return (uint64_t)e.x | ((uint64_t)e.y << 32);
    return ld_rec(R + q);
}
template <int CL>
__device__ __forceinline__ void fu_wait_rec(const uint64_t* R, const FuRingT<CL>& g, int q, uint32_t min_state) {
    uint32_t ns = FU_NS0;
    while (((uint32_t)fu_get<CL>(R, g, q) & 3u) < min_state) { __nanosleep(ns); if (ns < FU_NSMAX) ns += ns; }
}

// Decoupled look-back of unit u (u >= 1) over the chunk's records R[0..u).  Returns true with the number of
// symbols before the unit in `prefix`, or false with the true entry of the unit's first subsequence in
// `true_entry` when the unit has to re-walk (its guess differs from the FINAL exit of unit u - 1).
// Waiting (for a record to be published, or for the owner of a broken link to finalise) polls ONE record.
template <int CL>
__device__ bool fu_lookback(const uint64_t* __restrict__ R, const FuRingT<CL> ring_a, uint32_t u, uint32_t my_entry0, uint32_t lane,
                            uint32_t& prefix, uint32_t& true_entry) {
    uint32_t acc = 0, expect = my_entry0;
    int base = (int)u - 1;
    // (no separate wait for the predecessor's record: the window read below finds it EMPTY and polls it then - one
    //  round trip to L2 less per unit when the record comes from another CTA)
    for (;;) {
        const int j = base - (int)lane;
        const uint64_t rec = j >= 0 ? fu_get<CL>(R, ring_a, j) : 0ull;
        const uint32_t w = (uint32_t)rec;
        const uint32_t st = w & 3, en = (w >> 2) & 63, ex = (w >> 8) & 63, cnt = (w >> 14) & 0x3FFFF;
        const uint32_t en_up = __shfl_up_sync(0xffffffffu, en, 1);
        const uint32_t E = lane == 0 ? expect : en_up;            // entry used by the unit after unit j
        const uint32_t me = __ballot_sync(0xffffffffu, st == FU_EMPTY);
        const uint32_t mf = __ballot_sync(0xffffffffu, st == FU_FINAL);
        const uint32_t mb = __ballot_sync(0xffffffffu, st != FU_EMPTY && ex != E);
        const uint32_t any = me | mf | mb;
        if (!any) {                                               // 32 consistent speculative records: keep going back
            acc += warp_sum(cnt); expect = __shfl_sync(0xffffffffu, en, 31); base -= 32;
            continue;
        }
        const uint32_t d = __ffs(any) - 1;                        // first decisive record
        if ((mb >> d) & 1) {
            // the link between unit base - d and its successor is broken: the successor re-walks once its
            // predecessor is FINAL.  Mine: do that; somebody else's: wait until that unit has finalised.
            if (d == 0 && base == (int)u - 1) {
                if (mf & 1) { true_entry = __shfl_sync(0xffffffffu, ex, 0); return false; }
                fu_wait_rec<CL>(R, ring_a, base, FU_FINAL);
            } else {
                fu_wait_rec<CL>(R, ring_a, base - (int)d + 1, FU_FINAL);
            }
        } else if ((mf >> d) & 1) {
            prefix = __shfl_sync(0xffffffffu, (uint32_t)(rec >> 32), d) + acc + warp_sum(lane < d ? cnt : 0u);
            return true;
        } else {
            fu_wait_rec<CL>(R, ring_a, base - (int)d, FU_SPEC);        // not published yet
        }
        acc = 0; expect = my_entry0; base = (int)u - 1;           // start over
    }
}

// ---------------------------------------------------------------------------------------------
// stage geometry of one unit
// ---------------------------------------------------------------------------------------------
struct UnitGeom {
    uint64_t a0;            // 16-byte aligned global address of stage byte 0
    int32_t need;           // stage bytes the unit uses (multiple of 16)
    int32_t vlo, vhi;       // stage-relative byte range that belongs to the chunk
    int32_t tlo, thi;       // stage-relative byte range delivered by the bulk copy
    uint32_t bit0;          // stage-relative bit index of the unit's first bit (multiple of 8, >= 256)
};
__device__ __forceinline__ UnitGeom fu_geom(const uint8_t* comp, uint64_t comp_bytes, uint64_t chunk_off,
                                            uint32_t chunk_size, uint32_t u, uint32_t S) {
    UnitGeom g;
    const uint64_t cb = reinterpret_cast<uint64_t>(comp) + chunk_off;
    const uint64_t us = cb + (uint64_t)u * (128u * S);
    g.a0 = (us - 4 * FU_LEAD_MAX) & ~(uint64_t)15;
    g.bit0 = (uint32_t)(us - g.a0) * 8;
    g.need = (int32_t)(((us - g.a0) + 128u * S + 32 + 15) & ~(uint64_t)15);
    if (g.a0 >= cb && g.a0 + (uint64_t)g.need <= cb + chunk_size) {   // (warp-uniform) the stage lies inside the chunk:
        g.vlo = g.tlo = 0; g.vhi = g.thi = g.need;                    // all of it exists and is delivered by the bulk copy
        return g;
    }
    int64_t vlo = (int64_t)cb - (int64_t)g.a0, vhi = vlo + chunk_size;
    if (vlo < 0) vlo = 0;
    if (vhi > g.need) vhi = g.need;
    if (vhi < vlo) vhi = vlo;
    g.vlo = (int32_t)vlo; g.vhi = (int32_t)vhi;
    const uint64_t blo = (reinterpret_cast<uint64_t>(comp) + 15) & ~(uint64_t)15;
    const uint64_t bhi = (reinterpret_cast<uint64_t>(comp) + comp_bytes) & ~(uint64_t)15;
    uint64_t tl = g.a0 > blo ? g.a0 : blo;
    uint64_t th = g.a0 + g.need < bhi ? g.a0 + g.need : bhi;
    // nothing outside the chunk is needed: do not fetch it
    const uint64_t cl = (g.a0 + (uint64_t)g.vlo) & ~(uint64_t)15, ch = (g.a0 + (uint64_t)g.vhi + 15) & ~(uint64_t)15;
    if (tl < cl) tl = cl;
    if (th > ch) th = ch;
    if (th < tl) th = tl;
    g.tlo = (int32_t)(tl - g.a0); g.thi = (int32_t)(th - g.a0);
    return g;
}

// issued by ONE lane
__device__ __forceinline__ void fu_stage_issue(uint32_t stage_a, uint32_t bar_a, const UnitGeom& g) {
    const uint32_t bytes = (uint32_t)(g.thi - g.tlo);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (bytes) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(stage_a + (uint32_t)g.tlo), "l"(g.a0 + (uint64_t)g.tlo), "r"(bytes), "r"(bar_a) : "memory");
    } else {
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_a) : "memory");
    }
}
__device__ __forceinline__ void fu_mbar_wait(uint32_t bar_a, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "FW_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra FW_DONE;\n"
        "bra FW_LOOP;\n"
        "FW_DONE:\n"
        "}\n" ::"r"(bar_a), "r"(parity) : "memory");
}

// After the bulk copy landed.  The stage keeps the stream's bytes as they are (the bit reader byte-swaps a word when
// it becomes current, at no cost: the refill's register move is a PRMT instead), so nothing is left to do for a unit
// whose stage lies inside the chunk and was delivered whole.  Units at a chunk's ends: bytes outside the chunk read
// as zero (TableBasedHuffmanDecoder.java:204-208), chunk bytes the 16-byte aligned copy could not deliver are fetched
// one by one.  All lanes of the warp.
__device__ __forceinline__ void fu_stage_prepare(uint32_t stage_a, const UnitGeom& g, uint32_t lane) {
    const int32_t glo = g.vlo > g.tlo ? g.vlo : g.tlo, ghi = g.vhi < g.thi ? g.vhi : g.thi;   // good bytes
    if (glo == 0 && ghi >= g.need) return;
    for (int32_t b0 = (int32_t)lane * 16; b0 < g.need; b0 += 32 * 16) {
        if (b0 >= glo && b0 + 16 <= ghi) continue;
        if (b0 + 16 <= g.vlo || b0 >= g.vhi) { sts128(stage_a + b0, make_uint4(0u, 0u, 0u, 0u)); continue; }   // outside the chunk
        uint32_t x0 = 0, x1 = 0, x2 = 0, x3 = 0;
        for (int32_t j = 0; j < 16; ++j) {
            const int32_t b = b0 + j;
            uint32_t v = 0;
            if (b >= g.vlo && b < g.vhi)
                v = (b >= g.tlo && b < g.thi) ? lds8(stage_a + b) : *reinterpret_cast<const uint8_t*>(g.a0 + (uint64_t)b);
            v <<= 8 * (j & 3);
            if (j < 4) x0 |= v; else if (j < 8) x1 |= v; else if (j < 12) x2 |= v; else x3 |= v;
        }
        sts128(stage_a + b0, make_uint4(x0, x1, x2, x3));
    }
}

// ---------------------------------------------------------------------------------------------
// bit reader + walks.  Packed counter C: bits 16-27 stream position relative to a multiple of 32 at or
// before the walk's start (so (C >> 16) & 31 is the reader's funnel-shift amount), bits 0-15 output bits
// (8 per symbol) + FU_OUT_BIAS, whose low five bits are the shift that puts a lookup's symbols behind the
// bytes already in the accumulator.  One add of the table entry's .y advances both; bit 30 of an entry's .y
// (several candidate lengths / no code) makes the sum fail the loop's single compare; .y >= 13 << 16 marks a
// single-length long code (resolved by predicated instructions; the add is the same as for any entry).
// ---------------------------------------------------------------------------------------------
struct Reader { uint32_t hi, lo, nx, wa; };      // stream words j, j+1 (big-endian values), word j+2 as stored, its shared address
__device__ __forceinline__ void fu_seek(Reader& r, uint32_t stage_a, uint32_t pos) {
    const uint32_t a = stage_a + ((pos >> 5) << 2);
    r.hi = bswap32(lds32(a)); r.lo = bswap32(lds32(a + 4)); r.nx = lds32(a + 8); r.wa = a + 8;
}

// (the reader's pointer advances by a multiply-add of the toggled bit - 2^21 * 2^13 >> 32 = 4 - instead of a predicated
//  add, which ptxas turns into an add and a select)
#define FU_REFILL(CI, CO)                                                     \
    "lop3.b32 a, " CI ", " CO ", 0x200000, 0x28;\n"                           \
    "setp.ne.u32 p0, a, 0;\n"                                                 \
    "@p0 mov.u32 %1, %2;\n"                                                   \
    "@p0 prmt.b32 %2, %3, z, 0x0123;\n"                                       \
    "@p0 ld.shared.u32 %3, [%4+4];\n"                                         \
    "mad.hi.u32 %4, a, 8192, %4;\n"

// resolves an entry with several candidate lengths / no code: l = its length, 0 when nothing matches
#define FU_RARE_LEN(SFX, TAB, AUXOFF)                                         \
    "ld.shared.v2.u32 {ex, ey}, [ix];\n"                                      \
    "shr.u32 l, ex, 5;\n"                                                     \
    "and.b32 l, l, 63;\n"                                                     \
    "shr.u32 m, ex, 11;\n"                                                    \
    "and.b32 m, m, 63;\n"                                                     \
    "add.u32 auxb, " TAB ", " AUXOFF ";\n"                                    \
    HZ_PTX_LONGLEN(SFX)

// Walk WITHOUT output from C until the position field reaches Cend's: used for the lead-in before a
// subsequence and to finish a subsequence whose row is full.  Cb / exl = counter before and symbols of the
// last lookup (fu_settle() needs them).
__device__ __forceinline__ void fu_skim(uint32_t& C, Reader& r, uint32_t Cend, uint32_t wlut_a, uint32_t& Cb, uint32_t& exl) {
#define FU_SSTEP(CI, CO, SFX)                                                 \
    "shr.u32 s, " CI ", 16;\n"                                                \
    "shf.l.wrap.b32 v, %2, %1, s;\n"                                          \
    "shr.u32 ix, v, 20;\n"                                                    \
    "mad.lo.u32 ix, ix, 8, %7;\n"                                             \
    "ld.shared.v2.u32 {ex, ey}, [ix];\n"                                      \
    "add.u32 " CO ", " CI ", ey;\n"                                           \
    "setp.ge.u32 px, " CO ", %8;\n"                                           \
    "@px bra FS_CHECK" SFX ";\n"                                              \
    "FS_BACK" SFX ":\n"                                                       \
    FU_REFILL(CI, CO)
#define FU_SCHECK(CI, CO, SFX, FIN)                                           \
    "FS_CHECK" SFX ":\n"                                                      \
    "and.b32 t, " CO ", 0x40000000;\n"                                        \
    "setp.eq.u32 pq, t, 0;\n"                                                 \
    "@pq bra FS_LAST" SFX ";\n"                                               \
    FU_RARE_LEN("S" SFX, "%7", "%9")                                          \
    "max.u32 l, l, 1;\n"                                                      \
    "shl.b32 t, l, 16;\n"                                                     \
    "add.u32 " CO ", " CI ", t;\n"                                            \
    "add.u32 " CO ", " CO ", 8;\n"                                            \
    "mov.u32 ex, 0;\n"                                                        \
    "setp.lt.u32 pq, " CO ", %8;\n"                                           \
    "@pq bra FS_BACK" SFX ";\n"                                               \
    "FS_LAST" SFX ":\n"                                                       \
    FU_REFILL(CI, CO)                                                         \
    "mov.u32 %5, " CI ";\n"                                                   \
    "mov.u32 %6, ex;\n"                                                       \
    FIN                                                                       \
    "bra FS_DONE;\n"
    asm volatile(
        "{\n"
        ".reg .pred pl, px, p0, pq;\n"
        ".reg .u32 D, s, v, ix, ex, ey, t, u, a, z, l, m, auxb;\n"
        "mov.u32 z, 0;\n"
        "FS_TOP:\n"
        FU_SSTEP("%0", "D", "1")
        FU_SSTEP("D", "%0", "2")
        "bra FS_TOP;\n"
        FU_SCHECK("%0", "D", "1", "mov.u32 %0, D;\n")
        FU_SCHECK("D", "%0", "2", "")
        "FS_DONE:\n"
        "}\n"
        : "+r"(C), "+r"(r.hi), "+r"(r.lo), "+r"(r.nx), "+r"(r.wa), "=&r"(Cb), "=&r"(exl)
        : "r"(wlut_a), "r"(Cend), "n"(LUTN * 8)
        : "memory");
#undef FU_SSTEP
#undef FU_SCHECK
}

// Walk WITH output: symbols go to the lane's word-interleaved row (sp = shared address of the next word,
// acc = bytes not stored yet).  Stops when the position field reaches Cend's (ovf = 0) or, tested once per
// two lookups, when the row is full (ovf = 1; the caller finishes with fu_skim).  badc = smallest output
// field at which a bit pattern matched no codeword.
__device__ __forceinline__ void fu_walk(uint32_t& C, Reader& r, uint32_t Cend, uint32_t wlut_a, uint32_t& acc, uint32_t& sp,
                                        uint32_t& Cb, uint32_t& exl, uint32_t& badc, uint32_t& ovf) {
#define FU_COMMIT(CI, CO)                                                     \
    "lop3.b32 a, " CI ", " CO ", 0x200000, 0x28;\n"                           \
    "setp.ne.u32 p0, a, 0;\n"                                                 \
    "lop3.b32 u, " CI ", " CO ", 32, 0x28;\n"                                 \
    "setp.ne.u32 p1, u, 0;\n"                                                 \
    "@p0 mov.u32 %1, %2;\n"                                                   \
    "@p0 prmt.b32 %2, %3, z, 0x0123;\n"                                       \
    "@p0 ld.shared.u32 %3, [%4+4];\n"                                         \
    "mad.hi.u32 %4, a, 8192, %4;\n"                                           \
    "shf.l.wrap.b32 a, z, ex, " CI ";\n"                                      \
    "shf.l.wrap.b32 t, ex, z, " CI ";\n"                                      \
    "add.u32 %5, %5, a;\n"                                                    \
    "@p1 st.shared.u32 [%6], %5;\n"                                           \
    "mad.lo.u32 %6, u, 4, %6;\n"                                              \
    "selp.b32 %5, t, %5, p1;\n"
#define FU_WSTEP(CI, CO, SFX)                                                 \
    "shr.u32 s, " CI ", 16;\n"                                                \
    "shf.l.wrap.b32 v, %2, %1, s;\n"                                          \
    "shr.u32 ix, v, 20;\n"                                                    \
    "mad.lo.u32 ix, ix, 8, %11;\n"                                            \
    "ld.shared.v2.u32 {ex, ey}, [ix];\n"                                      \
    "shf.r.wrap.b32 t, v, z, ex;\n"                                           \
    "shr.s32 a, ex, 5;\n"                                                     \
    "add.u32 a, a, t;\n"                                                      \
    "add.u32 a, a, %11;\n"                                                    \
    "add.u32 " CO ", " CI ", ey;\n"                                           \
    "setp.ge.u32 pl, ey, %16;\n"                                                \
    "@pl ld.shared.u8 ex, [a];\n"                                             \
    "setp.ge.u32 px, " CO ", %12;\n"                                          \
    "@px bra FW_CHECK" SFX ";\n"                                              \
    "FW_BACK" SFX ":\n"                                                       \
    FU_COMMIT(CI, CO)
#define FU_WCHECK(CI, CO, SFX, FIN)                                           \
    "FW_CHECK" SFX ":\n"                                                      \
    "and.b32 t, " CO ", 0x40000000;\n"                                        \
    "setp.eq.u32 pq, t, 0;\n"                                                 \
    "@pq bra FW_LAST" SFX ";\n"                                               \
    FU_RARE_LEN("W" SFX, "%11", "%13")                                        \
    "setp.eq.u32 pq, l, 0;\n"                                                 \
    "@pq bra FW_BAD" SFX ";\n"                                                \
    "mad.lo.u32 a, l, 4, auxb;\n"                                             \
    "ld.shared.u32 t, [a+%14];\n"                                             \
    "sub.u32 u, 32, l;\n"                                                     \
    "shr.u32 u, v, u;\n"                                                      \
    "add.u32 a, t, u;\n"                                                      \
    "add.u32 a, a, auxb;\n"                                                   \
    "ld.shared.u8 ex, [a+%15];\n"                                             \
    "bra FW_GOT" SFX ";\n"                                                    \
    "FW_BAD" SFX ":\n"                                                        \
    "mov.u32 l, 1;\n"                                                         \
    "mov.u32 ex, 0;\n"                                                        \
    "and.b32 t, " CI ", 0xFFFF;\n"                                            \
    "min.u32 %7, %7, t;\n"                                                    \
    "FW_GOT" SFX ":\n"                                                        \
    "shl.b32 t, l, 16;\n"                                                     \
    "add.u32 " CO ", " CI ", t;\n"                                            \
    "add.u32 " CO ", " CO ", 8;\n"                                            \
    "setp.lt.u32 pq, " CO ", %12;\n"                                          \
    "@pq bra FW_BACK" SFX ";\n"                                               \
    "FW_LAST" SFX ":\n"                                                       \
    FU_COMMIT(CI, CO)                                                         \
    "mov.u32 %8, " CI ";\n"                                                   \
    "mov.u32 %9, ex;\n"                                                       \
    FIN                                                                       \
    "mov.u32 %10, 0;\n"                                                       \
    "bra FW_DONE;\n"
    asm volatile(
        "{\n"
        ".reg .pred pl, px, p0, p1, pq, po;\n"
        ".reg .u32 D, s, v, ix, ex, ey, t, u, a, z, l, m, auxb;\n"
        "mov.u32 z, 0;\n"
        "FW_TOP:\n"
        FU_WSTEP("%0", "D", "1")
        FU_WSTEP("D", "%0", "2")
        "and.b32 t, %0, 0x8000;\n"
        "setp.eq.u32 po, t, 0;\n"
        "@po bra FW_TOP;\n"
        "mov.u32 %10, 1;\n"
        "mov.u32 %8, %0;\n"
        "mov.u32 %9, 0;\n"
        "bra FW_DONE;\n"
        FU_WCHECK("%0", "D", "1", "mov.u32 %0, D;\n")
        FU_WCHECK("D", "%0", "2", "")
        "FW_DONE:\n"
        "}\n"
        : "+r"(C), "+r"(r.hi), "+r"(r.lo), "+r"(r.nx), "+r"(r.wa), "+r"(acc), "+r"(sp), "+r"(badc),
          "=&r"(Cb), "=&r"(exl), "=&r"(ovf)
        : "r"(wlut_a), "r"(Cend), "n"(LUTN * 8), "n"(offsetof(DecAux, symbase)), "n"(offsetof(DecAux, sorted)),
          "n"((LUTB + 1) << 16)
        : "memory");
#undef FU_WSTEP
#undef FU_WCHECK
#undef FU_COMMIT
}

// The walks stop at the first lookup that ENDS at or beyond the limit; that lookup may hold several symbols,
// some of which begin at or after the limit (they belong to the next subsequence).  Takes them back:
// returns the walk's counter with the position of the first codeword boundary >= limit and only the symbols
// that begin before it.  Cb / exl = counter before and symbols of the last lookup, lim = position field of the limit.
__device__ __forceinline__ uint32_t fu_settle(uint32_t C, uint32_t Cb, uint32_t exl, uint32_t lim, uint32_t aux_a) {
    const uint32_t n = ((C - Cb) & 0xFFFFu) >> 3;
    if (n > 1 && (C >> 16) > lim) {
        // straight-line (the lanes of a warp that get here differ in n): the first three symbols' lengths at once,
        // symbol k (k >= 1) stays iff it exists and the boundary before it lies before the limit
        const uint32_t la = aux_a + (uint32_t)offsetof(DecAux, len);
        const uint32_t l0 = lds8(la + (exl & 0xFFu)), l1 = lds8(la + ((exl >> 8) & 0xFFu)), l2 = lds8(la + ((exl >> 16) & 0xFFu));
        const uint32_t q0 = (Cb >> 16) + l0, q1 = q0 + l1, q2 = q1 + l2;
        const bool t1 = q0 < lim, t2 = t1 && n > 2 && q1 < lim, t3 = t2 && n > 3 && q2 < lim;
        const uint32_t j = 1u + t1 + t2 + t3;
        const uint32_t q = t3 ? (C >> 16) : (t2 ? q2 : (t1 ? q1 : q0));
        C = (q << 16) | ((Cb & 0xFFFFu) + 8 * j);
    }
    return C;
}

// exactly one codeword at the 32 stream bits v (slow paths): returns its length, the symbol in sym
__device__ __forceinline__ uint32_t fu_one(const DecAux& A, uint32_t wlut_a, uint32_t v, uint32_t& sym, bool& bad) {
    const uint2 e = lds64(wlut_a + ((v >> 20) << 3));
    if (e.y < ((LUTB + 1u) << 16)) { sym = e.x & 0xFFu; return A.len[sym]; }
    const uint32_t l = long_len(A, v, LUTB + 1, (uint32_t)A.maxlen);
    if (!l) { bad = true; sym = 0; return 1; }
    sym = A.sorted[A.symbase[l] + (int32_t)(v >> (32 - l))];
    return l;
}

// ---------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------
// -DFU_CHECK (developer build; compute-sanitizer is closed on the B200 pool): bounds assertions of our own on every
// shared-memory region and global range the kernel addresses; a violation prints once per thread and latches HZ_ERR_CUDA
#ifdef FU_CHECK
#define FU_ASSERT(cond, what) do { if (!(cond)) { printf("FU_CHECK %s failed: block %d thread %d\n", what, (int)blockIdx.x, (int)threadIdx.x); hz_set_status(a.status, HZ_ERR_CUDA); } } while (0)
#else
#define FU_ASSERT(cond, what) do { } while (0)
#endif
// -DFU_TIMING (developer build): per-phase clock64 totals over all warps, printed by the launcher
#ifdef FU_TIMING
#define FU_T(i) do { const long long now__ = clock64(); tim__[i] += now__ - last__; last__ = now__; } while (0)
// per-warp event trace of CTA 0 (a.tim[16] = count, a.tim[32 + 2 i] = clock, [33 + 2 i] = warp << 48 | event << 32 | value)
#define FU_TRACE(ev, val) do { if (blockIdx.x == 0 && lane == 0 && a.tim) { const unsigned long long ix__ = atomicAdd(a.tim + 16, 1ull); \
    if (ix__ < 4000) { a.tim[32 + 2 * ix__] = (unsigned long long)clock64(); a.tim[33 + 2 * ix__] = ((unsigned long long)wid << 48) | ((unsigned long long)(ev) << 32) | (unsigned long long)(uint32_t)(val); } } } while (0)
#else
#define FU_T(i) do { } while (0)
#define FU_TRACE(ev, val) do { } while (0)
#endif
struct FuArgs {
    const uint8_t* comp; uint64_t comp_bytes;
    const uint64_t* comp_off; const uint32_t* comp_size; const uint32_t* orig_size;
    uint32_t K; FuPlan P; const uint8_t* tables; uint64_t* rec;
    uint8_t* out; uint64_t out_cap; int* status;
    unsigned long long* tim;   // -DFU_TIMING
    uint32_t* dbg;          // developer dump (HZ_FU_DUMP): entry | exit << 8 | count << 16 per subsequence
};

// Chunk for this CTA.  Thread 0 holds a ticket drawn one chunk AHEAD (`ahead`: the atomic's round trip ran under
// the previous chunk's decode), resolves it to a chunk that has units, requests that chunk's table (ONE bulk copy
// of wlut + aux, completion on S.tbar) and draws the next ticket.  When the tickets have run out (few large
// chunks) the CTA joins a chunk that still has units.  Called by all threads after a barrier that follows the last
// use of the previous table.
template <class SH>
__device__ __forceinline__ void fu_table_issue(const FuArgs& a, SH& S, uint32_t k) {
    const uint32_t bar_a = smem_u32(&S.tbar), dst = smem_u32(S.wlut);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "n"(FU_TABLE_BYTES) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(a.tables + (size_t)k * FU_TABLE_BYTES), "n"(FU_TABLE_BYTES), "r"(bar_a) : "memory");
}
// cluster pair (CL == 2): rank of this CTA and DSMEM helpers
__device__ __forceinline__ uint32_t fu_cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t fu_mapa(uint32_t local_a, uint32_t rank) {
    uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_a), "r"(rank)); return r;
}
__device__ __forceinline__ void fu_remote_st32(uint32_t ra, uint32_t v) {
    asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(ra), "r"(v) : "memory");
}
__device__ __forceinline__ void fu_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// CL == 2: the pair works as ONE team - the even CTA picks (ticket, then helping), the odd CTA FOLLOWS it: it waits for
// the peer's announcement (S.peer_k of this CTA, written by the peer through DSMEM; bounded wait, then it picks for
// itself) and takes the same chunk, so that the chunk's units are shared by two CTAs whose rings see each other's
// records.  Every CTA announces the chunk it is on to its peer (0xFFFFFFFF: no more work); stale values only cost a
// missed or a useless push: ring entries carry their chunk, and every record is in global memory as well.
// `ahead` of the odd CTA holds the last announcement it followed.
#define FU_ANN_NONE 0xFFFFFFFFu
template <class SH, int CL>
__device__ uint32_t fu_pick_chunk(const FuArgs& a, SH& S, uint32_t& ahead, uint32_t cl_rank) {
    if constexpr (CL != 2) {                              // a lone CTA: ticket drawn ahead, then helping
        const uint32_t K = a.K;
        if (threadIdx.x == 0) {
            uint32_t k = FU_NONE, t = ahead;
            for (;;) {
                if (t >= K) break;
                if (a.P.nunit[t]) { k = t; break; }
                t = atomicAdd(&a.P.ctl[0], 1u);
            }
            if (k != FU_NONE) { fu_table_issue(a, S, k); ahead = atomicAdd(&a.P.ctl[0], 1u); }
            else ahead = K;
            S.s_k = k; S.s_pick = FU_NONE;
        }
        if (threadIdx.x < FU_RING) S.ring[threadIdx.x] = make_uint4(0u, 0u, 0u, 0u);
        __syncthreads();
        if (S.s_k != FU_NONE) return S.s_k;
        if (K > 2048) return FU_NONE;                         // thousands of chunks balance by themselves
        // helping: the first chunk (from a start that spreads the CTAs) whose unit counter has not run out
        const uint32_t start = (uint32_t)(((uint64_t)blockIdx.x * K) / gridDim.x);
        for (uint32_t i0 = 0; i0 < K; i0 += blockDim.x) {
            const uint32_t i = i0 + threadIdx.x;
            if (i < K) {
                const uint32_t k = (start + i) % K;
                const uint32_t nu = a.P.nunit[k];
                if (nu && *reinterpret_cast<volatile uint32_t*>(a.P.unit_ctr + k) < nu) atomicMin(&S.s_pick, i);
            }
            __syncthreads();
            if (S.s_pick != FU_NONE) break;
        }
        __syncthreads();
        const uint32_t p = S.s_pick;
        if (p == FU_NONE) return FU_NONE;
        const uint32_t k = (start + p) % K;
        if (threadIdx.x == 0) fu_table_issue(a, S, k);
        return k;
    } else {
    const uint32_t K = a.K;
    bool follow = false;
    if (threadIdx.x == 0) {
        uint32_t k = FU_NONE;
        bool decided = false;
        if (CL == 2 && cl_rank == 1) {
            uint32_t pk = ahead, spins = 0;
            while ((pk = *fu_peer_k(S)) == ahead && ++spins < (1u << 16)) __nanosleep(64);
            if (pk != ahead) {
                ahead = pk; decided = true;
                if (pk != FU_ANN_NONE && pk != 0) { k = pk - 1; fu_table_issue(a, S, k); }
            }
        }
        if (!decided) {
            uint32_t t = (CL == 2 && cl_rank == 1) ? atomicAdd(&a.P.ctl[0], 1u) : ahead;   // (a follower that timed out holds no ticket)
            for (;;) {
                if (t >= K) break;
                if (a.P.nunit[t]) { k = t; break; }
                t = atomicAdd(&a.P.ctl[0], 1u);
            }
            if (k != FU_NONE) { fu_table_issue(a, S, k); if (!(CL == 2 && cl_rank == 1)) ahead = atomicAdd(&a.P.ctl[0], 1u); }
            else if (!(CL == 2 && cl_rank == 1)) ahead = K;
        }
        S.s_k = k; S.s_pick = (CL == 2 && cl_rank == 1 && decided) ? 0u : FU_NONE;    // s_pick == 0: the follower does not search
    }
    for (uint32_t i = threadIdx.x; i < (CL == 2 ? FU_RING_CL : FU_RING); i += blockDim.x) S.ring[i] = make_uint4(0u, 0u, 0u, 0u);
    __syncthreads();
    uint32_t k = S.s_k;
    if (CL == 2) {
        follow = cl_rank == 1 && S.s_pick == 0u;
        __syncthreads();
        if (threadIdx.x == 0 && follow) S.s_pick = FU_NONE;
        __syncthreads();
    }
    if (k == FU_NONE && K <= 2048 && !follow) {           // (thousands of chunks balance by themselves)
        // helping: the first chunk (from a start that spreads the CTAs) whose unit counter has not run out
        const uint32_t start = (uint32_t)(((uint64_t)blockIdx.x * K) / gridDim.x);
        for (uint32_t i0 = 0; i0 < K; i0 += blockDim.x) {
            const uint32_t i = i0 + threadIdx.x;
            if (i < K) {
                const uint32_t c = (start + i) % K;
                const uint32_t nu = a.P.nunit[c];
                if (nu && *reinterpret_cast<volatile uint32_t*>(a.P.unit_ctr + c) < nu) atomicMin(&S.s_pick, i);
            }
            __syncthreads();
            if (S.s_pick != FU_NONE) break;
        }
        __syncthreads();
        const uint32_t p = S.s_pick;
        if (p != FU_NONE) {
            k = (start + p) % K;
            if (threadIdx.x == 0) fu_table_issue(a, S, k);
        }
    }
    if (CL == 2 && threadIdx.x == 0)
        fu_remote_st32(fu_mapa(smem_u32((const void*)fu_peer_k(S)), cl_rank ^ 1u), k == FU_NONE ? FU_ANN_NONE : k + 1);
    return k;
    }
}

// W = warps per CTA, MINB = CTAs per SM: <24, 1> for chunks of many units; <8, 2> and <5, 3> for streams of small
// chunks, where several independent CTAs per SM (each with its own chunk's table) overlap one CTA's per-chunk
// serial part (ticket, table copy, barriers, the chain of a few units) with the others' walks
// CL = 2: launched as cluster pairs (streams with fewer chunks than SMs): the pair shares chunks through its rings.
template <int W, int MINB, int CL>
__global__ void __launch_bounds__(W * 32, MINB)
dec_fused_kernel(const FuArgs a) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    typedef FuSharedT<CL == 2 ? FU_RING_CL : FU_RING> SH;
    SH& S = *reinterpret_cast<SH*>(smem_raw);
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t stage_a = pin_reg(smem_u32(smem_raw + FU_SHARED_BYTES_RN(CL == 2 ? FU_RING_CL : FU_RING) + (size_t)wid * FU_WARP_BYTES));
    const uint32_t rows_a = stage_a + FU_STAGE_BYTES;
    const uint32_t bar_a = smem_u32(&S.bar[wid]);
    const uint32_t wlut_a = pin_reg(smem_u32(S.wlut)), aux_a = pin_reg(smem_u32(S.aux));
    FuRingT<CL> ring_a;
    ring_a.a = smem_u32(S.ring);
    if constexpr (CL == 2) { ring_a.ktag = 0; ring_a.peer = 0; }
    const uint32_t cl_rank = CL == 2 ? fu_cluster_rank() : 0u;
    const uint32_t peer_ring = CL == 2 ? fu_mapa(ring_a.a, cl_rank ^ 1u) : 0u;
    const DecAux& A = *reinterpret_cast<const DecAux*>(S.aux);
    if (lane == 0) {
        mbar_init(&S.bar[wid], 1);
        if (wid == 0) { mbar_init(&S.tbar, 1); if (CL == 2) *fu_peer_k(S) = 0; }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (CL == 2) fu_cluster_sync();                       // the peer CTA has started: its shared memory may be written
#ifdef FU_TIMING
    long long tim__[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, last__ = clock64();
#endif
    uint32_t phase = 0;                                   // parity of this warp's mbarrier
    uint32_t tphase = 0;                                  // parity of the table's mbarrier
    bool out_pending = false;                             // a bulk copy out of this warp's rows may still be reading them
    uint32_t ahead = 0;                                   // thread 0: the chunk ticket drawn ahead
    if (threadIdx.x == 0 && !(CL == 2 && cl_rank == 1)) ahead = atomicAdd(&a.P.ctl[0], 1u);   // (the odd CTA of a pair follows its peer)
    __syncthreads();

    for (;;) {
        const uint32_t k = fu_pick_chunk<SH, CL>(a, S, ahead, cl_rank);
        if (k == FU_NONE) break;
        if constexpr (CL == 2) ring_a.ktag = k + 1;       // (a lone CTA's ring is reset per chunk and only it writes to it)
        FU_T(11);
        FU_TRACE(0, k);
        // the chunk's geometry and this warp's first unit are fetched while the table is on its way
        const uint32_t osize = a.orig_size[k];
        const uint64_t ooff = a.P.orig_off[k];
        const uint32_t nunit = a.P.nunit[k];
        const uint64_t coff = a.comp_off[k];
        const uint32_t csize = a.comp_size[k];
        const uint32_t nsub = a.P.nsub[k];
        const uint32_t Sw = a.P.geom[k] & 0xFF, lead = a.P.geom[k] >> 8;
        const uint32_t sub_bits = Sw * 32;
        uint64_t* R = a.rec + a.P.unit_base[k];
        const uint64_t gout = reinterpret_cast<uint64_t>(a.out) + ooff;
        uint32_t u = 0;
        if (lane == 0) u = atomicAdd(a.P.unit_ctr + k, 1u);
        u = __shfl_sync(0xffffffffu, u, 0);
        if constexpr (CL == 2) ring_a.peer = *fu_peer_k(S) == k + 1 ? peer_ring : 0u;
        if (u < nunit && lane == 0) { fu_mark_pending<CL>(ring_a, u); fu_stage_issue(stage_a, bar_a, fu_geom(a.comp, a.comp_bytes, coff, csize, u, Sw)); }
        FU_T(8);
        fu_mbar_wait(smem_u32(&S.tbar), tphase); tphase ^= 1;
        FU_T(9);
        int reject = 0;
        if (A.bad) reject = HZ_ERR_BAD_LENGTHS;
        else if (ooff + osize > a.out_cap) reject = HZ_ERR_OUT_TOO_SMALL;
        if (reject) {
            // (the units already ticketed have their stage copies in flight: drain them, the stage is reused)
            if (u < nunit) { fu_mbar_wait(bar_a, phase); phase ^= 1; }
            if (threadIdx.x == 0) { hz_set_status(a.status, reject); atomicMax(a.P.unit_ctr + k, nunit); }
            __syncthreads();
            continue;
        }
        const uint32_t U = (uint32_t)A.uniform;

        // ---- warp loop over the chunk's units -------------------------------------------------
        while (u < nunit) {
            const UnitGeom g = fu_geom(a.comp, a.comp_bytes, coff, csize, u, Sw);
            FU_ASSERT(g.need > 0 && g.need <= FU_STAGE_BYTES && g.tlo >= 0 && g.tlo <= g.thi && g.thi <= g.need && (g.tlo & 15) == 0 && (g.thi & 15) == 0, "stage geometry");
            FU_ASSERT(g.a0 + (uint64_t)g.tlo >= (reinterpret_cast<uint64_t>(a.comp) & ~15ull) && g.a0 + (uint64_t)g.thi <= reinterpret_cast<uint64_t>(a.comp) + a.comp_bytes, "bulk copy inside the stream");
            FU_ASSERT(g.bit0 >= lead * 32 && g.bit0 / 8 + 128 * Sw + 32 <= (uint32_t)g.need + 15, "unit inside the stage");
            FU_T(0);
            FU_TRACE(1, u);
            fu_mbar_wait(bar_a, phase); phase ^= 1;
            FU_T(1);
            FU_TRACE(2, u);
            fu_stage_prepare(stage_a, g, lane);
            if (out_pending) {                            // the previous unit's window must have left the rows
                if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                out_pending = false;
            }
            __syncwarp();

            const uint32_t i = u * 32 + lane;             // subsequence index within the chunk
            const bool active = i < nsub;
            // stage-relative bits.  The chunk's last subsequence ends with the chunk: the zero bits behind it would decode
            // to a long run of the all-zero codeword's symbol (a row overflow in every chunk's last unit); symbols that
            // the stream does not hold are written by the tail fill below
            const uint32_t nominal = g.bit0 + lane * sub_bits;
            const uint64_t left = (uint64_t)csize * 8 - (uint64_t)i * sub_bits;      // (only meaningful for active lanes)
            const uint32_t end = nominal + (active && left < sub_bits ? (uint32_t)left : sub_bits);
            uint32_t entry = 0;
            if (active && i != 0) {
                if (U) {                                  // equal-length code: boundaries are the multiples of U
                    const uint64_t nomc = (uint64_t)i * sub_bits;
                    entry = (uint32_t)((U - nomc % U) % U);
                } else {                                  // guess: walk the lead-in, take the first boundary inside
                    const uint32_t p0 = nominal - lead * 32, org = p0 & ~31u;
                    Reader r; fu_seek(r, stage_a, p0);
                    uint32_t C = ((p0 - org) << 16) | FU_OUT_BIAS, Cb, exl;
                    fu_skim(C, r, (nominal - org) << 16, wlut_a, Cb, exl);
                    C = fu_settle(C, Cb, exl, nominal - org, aux_a);
                    entry = (C >> 16) - (nominal - org);
                }
            }
            FU_T(2);
            FU_TRACE(3, u);
            // decode (and re-decode where a guess was wrong) until the chain of the unit is consistent and
            // anchored in the chunk's earlier units
            bool need = active;
            uint32_t count = 0, exitv = 0, badc = 0xFFFFFFFFu, prefix = 0;
            bool published = false;
            for (;;) {
                if (need) {
                    const uint32_t p0 = nominal + entry, org = p0 & ~31u;
                    Reader r; fu_seek(r, stage_a, p0);
                    uint32_t C = ((p0 - org) << 16) | FU_OUT_BIAS, Cb, exl, ovf, acc = 0, sp = rows_a + lane * 4;
                    const uint32_t Cend = (end - org) << 16;
                    badc = 0xFFFFFFFFu;
                    FU_T(7);
                    fu_walk(C, r, Cend, wlut_a, acc, sp, Cb, exl, badc, ovf);
                    FU_T(3);
                    FU_TRACE(4, u);
                    FU_ASSERT(sp >= rows_a + lane * 4 && sp <= rows_a + lane * 4 + (FU_ROW_WORDS - 1) * 128, "row pointer");
                    FU_ASSERT(r.wa >= stage_a && r.wa + 4 < stage_a + (uint32_t)g.need, "reader inside the stage");
                    if (ovf) fu_skim(C, r, Cend, wlut_a, Cb, exl);
                    else sts32(sp, acc);                  // bytes still in the accumulator (the row has two guard words)
                    C = fu_settle(C, Cb, exl, end - org, aux_a);
                    count = ((C & 0xFFFFu) - FU_OUT_BIAS) >> 3;
                    exitv = (C >> 16) - (end - org);
                }
                const uint32_t want = __shfl_up_sync(0xffffffffu, exitv, 1);
                need = active && lane > 0 && want != entry;
                if (need) entry = want;
                if (__any_sync(0xffffffffu, need)) { published = false; continue; }
                // the last ACTIVE lane's exit is the unit's exit
                const uint32_t nact = min(32u, nsub - u * 32);
                const uint32_t uexit = __shfl_sync(0xffffffffu, exitv, nact - 1);
                const uint32_t uentry = __shfl_sync(0xffffffffu, entry, 0);
                const uint32_t ucount = warp_sum(active ? count : 0);
                if (u == 0) {
                    if (lane == 0) fu_put<CL>(R, ring_a, 0, fu_pack(FU_FINAL, 0, uexit, ucount, ucount));
                    prefix = 0;
                    break;
                }
                if (!published) {
                    if (lane == 0) fu_put<CL>(R, ring_a, u, fu_pack(FU_SPEC, uentry, uexit, ucount, 0));
                    published = true;
                }
                uint32_t true_entry = 0;
                FU_T(7);
                const bool lb_ok = fu_lookback<CL>(R, ring_a, u, uentry, lane, prefix, true_entry);
                FU_T(4);
                FU_TRACE(5, u);
                if (lb_ok) {
                    if (lane == 0) fu_put<CL>(R, ring_a, u, fu_pack(FU_FINAL, uentry, uexit, ucount, prefix + ucount));
                    break;
                }
                need = lane == 0;                         // the guess of the first subsequence was wrong: re-walk from the truth
                if (need) entry = true_entry;
                published = false;
            }
            __syncwarp();
            if (a.dbg) a.dbg[(size_t)(a.P.unit_base[k] + u) * 32 + lane] = entry | (exitv << 8) | (count << 16);
            FU_T(7);
            const uint32_t u_cur = u;
            FU_T(5);
            // ---- output ------------------------------------------------------------------------
            if (!active) count = 0;
            const uint32_t rel = warp_excl_scan(count, lane);
            const uint32_t obase = prefix + rel;                           // chunk-relative index of the lane's first symbol
            const uint32_t cw = obase >= osize ? 0u : min(count, osize - obase);   // symbols of this lane that exist
            const bool over = __any_sync(0xffffffffu, count > FU_CAP_SYMS);
            uint32_t tk = 0; bool tk_drawn = false;                            // the next unit's ticket (lane 0)
            if (badc != 0xFFFFFFFFu && obase + ((badc - FU_OUT_BIAS) >> 3) < osize) hz_set_status(a.status, HZ_ERR_DECODE);
            if (over) {
                // a row overflowed (far more symbols than the chunk's average in one subsequence): every lane
                // decodes its subsequence again, one codeword at a time, straight to global memory
                if (cw) {
                    uint32_t p = nominal + entry, idx = obase;
                    const uint32_t stop = obase + cw;
                    bool bad = false;
                    while (idx < stop) {
                        const uint32_t wa = stage_a + ((p >> 5) << 2);
                        const uint32_t v = __funnelshift_l(bswap32(lds32(wa + 4)), bswap32(lds32(wa)), p);
                        uint32_t sym;
                        p += fu_one(A, wlut_a, v, sym, bad);
                        *reinterpret_cast<uint8_t*>(gout + idx) = (uint8_t)sym;
                        ++idx;
                    }
                    if (bad) hz_set_status(a.status, HZ_ERR_DECODE);
                }
                __syncwarp();
            } else {
                const uint32_t T = warp_sum(cw);
                FU_ASSERT(T == 0 || (T <= 32 * FU_CAP_SYMS && (uint64_t)prefix + T <= osize && ooff + prefix + T <= a.out_cap), "window inside the output");
                if (T) {
                    const uint32_t first = __shfl_sync(0xffffffffu, obase, 0);
                    const uint64_t g0 = gout + first;                     // global address of the window's first symbol
                    const uint32_t shift = (uint32_t)(g0 & 15);
                    const uint32_t d = shift + (obase - first);           // window byte of the lane's first symbol
                    // all rows to registers (the window overlays the rows)
                    uint32_t rw[FU_ROW_WORDS - 1];
#pragma unroll
                    for (int j = 0; j < FU_ROW_WORDS - 1; ++j) rw[j] = lds32(rows_a + (j * 32 + lane) * 4);
                    const uint32_t hn = min(cw, (4u - (d & 3)) & 3);      // bytes before the lane's first whole window word
                    const uint32_t nfull = (cw - hn) >> 2, tn = (cw - hn) & 3;
                    uint32_t tail = 0;                                    // the last tn bytes (row bytes hn + 4 nfull ...)
                    if (tn) {                                             // they sit in ONE row word: bytes tb & 3 ... of word tb >> 2
                        const uint32_t tb = hn + 4 * nfull;
                        const uint32_t w = lds32(rows_a + ((tb >> 2) * 32 + lane) * 4), w2 = lds32(rows_a + (((tb >> 2) + 1) * 32 + lane) * 4);
                        tail = __funnelshift_r(w, w2, 8 * (tb & 3));
                    }
                    __syncwarp();
                    // the rows are in registers: draw the next ticket now, its round trip runs under the stores below
                    // (about a third of a microsecond before the last possible moment, see the note at the loop's end)
                    if (lane == 0) tk = atomicAdd(a.P.unit_ctr + k, 1u);
                    tk_drawn = true;
                    FU_ASSERT((cw == 0 || d + cw <= 15 + T) && 15 + T + 4 <= FU_ROWS_BYTES, "window inside the rows");
                    const uint32_t bs = (d & 3) * 8;
                    const uint32_t w0 = rows_a + (d & ~3u);               // window word that holds the first symbol
                    if (hn > 0) sts8(w0 + (d & 3), rw[0]);
                    if (hn > 1) sts8(w0 + (d & 3) + 1, rw[0] >> 8);
                    if (hn > 2) sts8(w0 + (d & 3) + 2, rw[0] >> 16);
                    const uint32_t jmin = bs ? 1u : 0u;
                    {
                        // window word w0 + 4 j = row bytes [4 j - (d & 3), +4)
                        uint32_t prev = 0;
#pragma unroll
                        for (int j = 0; j < FU_ROW_WORDS - 1; ++j) {
                            const uint32_t val = __funnelshift_l(prev, rw[j], bs);
                            if ((uint32_t)j >= jmin && (uint32_t)j < jmin + nfull) sts32(w0 + 4 * j, val);
                            prev = rw[j];
                        }
                    }
                    {
                        const uint32_t ta = w0 + 4 * (jmin + nfull);
                        if (tn > 0) sts8(ta, tail);
                        if (tn > 1) sts8(ta + 1, tail >> 8);
                        if (tn > 2) sts8(ta + 2, tail >> 16);
                    }
                    __syncwarp();
                    // window bytes [shift, shift + T) -> global [g0, g0 + T): whole 16-byte units by one bulk copy
                    const uint64_t gw = g0 - shift;                       // global address of window byte 0 (16-byte aligned)
                    const uint32_t b0 = shift, b1 = shift + T;
                    const uint32_t f0 = (b0 + 15) >> 4, f1 = b1 >> 4;
                    if (f1 > f0) {
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                        __syncwarp();
                        if (lane == 0) {
                            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                                         ::"l"(gw + (uint64_t)f0 * 16), "r"(rows_a + f0 * 16), "r"((f1 - f0) * 16) : "memory");
                            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                        }
                        out_pending = true;
                        // ragged ends, one byte per lane (lanes 0-15: first unit, lanes 16-31: last unit)
                        const uint32_t bb = lane < 16 ? (b0 & ~15u) + lane : (b1 & ~15u) + (lane - 16);
                        const bool edge = lane < 16 ? (b0 & 15) != 0 : (b1 & 15) != 0;
                        if (edge && bb >= b0 && bb < b1) *reinterpret_cast<uint8_t*>(gw + bb) = (uint8_t)lds8(rows_a + bb);
                    } else {
                        for (uint32_t bb = b0 + lane; bb < b1; bb += 32) *reinterpret_cast<uint8_t*>(gw + bb) = (uint8_t)lds8(rows_a + bb);
                    }
                }
            }
            FU_T(6);
            FU_TRACE(6, u_cur);
            // The next unit's ticket is drawn as LATE as possible, right before its bytes are fetched: a unit waits (in
            // its look-back) for every unit with a smaller ticket, so whatever a warp does between drawing a ticket and
            // publishing that unit's record delays all the warps behind it.  Drawing the ticket before the output
            // phase (to hide the fetch behind it) measured 1.4 % slower, drawing it a whole unit ahead 33 % slower, and handing
            // tickets out of a per-CTA pool of blocks drawn ahead 5 % slower (30 % where CTAs share a chunk).
            if (!tk_drawn && lane == 0) tk = atomicAdd(a.P.unit_ctr + k, 1u);
            u = __shfl_sync(0xffffffffu, tk, 0);
            if constexpr (CL == 2) ring_a.peer = *fu_peer_k(S) == k + 1 ? peer_ring : 0u;
            if (u < nunit && lane == 0) { fu_mark_pending<CL>(ring_a, u); fu_stage_issue(stage_a, bar_a, fu_geom(a.comp, a.comp_bytes, coff, csize, u, Sw)); }
            // the chunk's last unit: when the stream holds fewer symbols than orig_size the decoder goes on reading
            // zero bits (TableBasedHuffmanDecoder.java:204-208), i.e. the all-zero codeword's symbol repeats
            if (u_cur == nunit - 1) {
                const uint32_t have = prefix + __shfl_sync(0xffffffffu, rel + count, 31);
                if (have < osize) {
                    const uint8_t s0 = A.sorted[0];
                    for (uint64_t b = (uint64_t)have + lane; b < osize; b += 32) *reinterpret_cast<uint8_t*>(gout + b) = s0;
                }
            }
            __syncwarp();
        }
        FU_T(0);
        __syncthreads();                                  // every warp is done with this chunk's table
        FU_T(10);
        FU_TRACE(7, k);
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // shared memory must outlive the copies
    if (CL == 2) { __syncthreads(); fu_cluster_sync(); }  // ... and the peer's pushes into it: the pair leaves together
#ifdef FU_TIMING
    FU_T(0);
    if (lane == 0 && a.tim) for (int i = 0; i < 12; ++i) atomicAdd(a.tim + i, (unsigned long long)tim__[i]);
#endif
}

// ---------------------------------------------------------------------------------------------
// launcher.  `ident` = per-chunk identity flags (hz_decode.cu); identity chunks are copied by the caller with the
// slice prefix sums this plan provides.
// ---------------------------------------------------------------------------------------------
int hzk_decode_fused(hz_ctx* ctx, const uint8_t* d_comp, uint64_t comp_bytes, const uint64_t* d_comp_off,
                     const uint32_t* d_comp_size, const uint32_t* d_orig_size, const uint64_t* d_orig_off,
                     const uint8_t* d_len, uint32_t K, uint8_t* d_out, uint64_t out_cap, const uint8_t* d_ident,
                     const uint64_t** plan_orig_off, const uint32_t** plan_islice) {
    HZ_TRY(hz_reserve(ctx, &ctx->dec_meta, ((size_t)K + 1) * (6 * sizeof(uint32_t) + sizeof(uint64_t)) + 256));
    FuPlan P;
    uint8_t* m = (uint8_t*)ctx->dec_meta.p;
    P.orig_off = (uint64_t*)m; m += ((size_t)K + 1) * sizeof(uint64_t);
    P.nsub = (uint32_t*)m; m += ((size_t)K + 1) * sizeof(uint32_t);
    P.nunit = (uint32_t*)m; m += ((size_t)K + 1) * sizeof(uint32_t);
    P.unit_base = (uint32_t*)m; m += ((size_t)K + 1) * sizeof(uint32_t);
    P.geom = (uint32_t*)m; m += ((size_t)K + 1) * sizeof(uint32_t);
    P.unit_ctr = (uint32_t*)m; m += ((size_t)K + 1) * sizeof(uint32_t);
    P.islice = (uint32_t*)m; m += ((size_t)K + 1) * sizeof(uint32_t);
    P.ctl = (uint32_t*)m;
    // every chunk has at most ceil(comp_size / (4 * FU_SUB_MIN * 32)) + 1 units
    const uint64_t max_units = comp_bytes / (128ull * FU_SUB_MIN) + 2ull * K + 2;
    if (max_units > 0xFFFFFFFFull) return hz_fail(ctx, HZ_ERR_ARG, "decode: stream too large for one call");
    HZ_TRY(hz_reserve(ctx, &ctx->dec_rec, max_units * sizeof(uint64_t)));
    HZ_TRY(hz_reserve(ctx, &ctx->dec_tables, (size_t)K * FU_TABLE_BYTES));
    if (!ctx->attr_decode_fused) {
        HZ_CUDA(ctx, cudaFuncSetAttribute(dec_fused_kernel<24, 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FU_SMEM_BYTES(24)));
        HZ_CUDA(ctx, cudaFuncSetAttribute(dec_fused_kernel<24, 1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FU_SMEM_BYTES_CL(24, 2)));
        HZ_CUDA(ctx, cudaFuncSetAttribute(dec_fused_kernel<8, 2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FU_SMEM_BYTES(8)));
        HZ_CUDA(ctx, cudaFuncSetAttribute(dec_fused_kernel<5, 3, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FU_SMEM_BYTES(5)));
        ctx->attr_decode_fused = true;
    }
    const uint32_t lead_knob = (uint32_t)ctx->knobs.fu_lead;   // developer knob
    HZ_LAUNCH(ctx, "dec_plan", fu_plan_kernel, 1, 1024, 0, d_comp_off, d_comp_size, d_orig_size, d_orig_off, comp_bytes, K, P,
              d_ident, lead_knob, ctx->d_status);
    HZ_LAUNCH(ctx, "dec_zero", fu_zero_kernel, 2 * ctx->sm_count, 256, 0, (uint64_t*)ctx->dec_rec.p, P.unit_base, K);
    HZ_LAUNCH(ctx, "dec_tables", fu_tables_kernel, K, DT, 0, d_len, P, (uint8_t*)ctx->dec_tables.p);
    FuArgs a;
    a.comp = d_comp; a.comp_bytes = comp_bytes; a.comp_off = d_comp_off; a.comp_size = d_comp_size; a.orig_size = d_orig_size;
    a.K = K; a.P = P; a.tables = (const uint8_t*)ctx->dec_tables.p; a.rec = (uint64_t*)ctx->dec_rec.p;
    a.out = d_out; a.out_cap = out_cap; a.status = ctx->d_status; a.dbg = nullptr; a.tim = nullptr;
#ifdef FU_TIMING
    cudaMallocManaged(&a.tim, (32 + 8000) * sizeof(unsigned long long)); cudaMemset(a.tim, 0, (32 + 8000) * sizeof(unsigned long long));
#endif
    const char* dump = ctx->knobs.fu_dump.empty() ? nullptr : ctx->knobs.fu_dump.c_str();   // developer knob: per-subsequence records to a file
    if (dump) { cudaMallocManaged(&a.dbg, max_units * 32 * sizeof(uint32_t)); cudaMemset(a.dbg, 0xFF, max_units * 32 * sizeof(uint32_t)); }
    // CTA shape by the stream's average chunk (no host synchronisation: from the call's sizes): units per chunk
    uint32_t warps = 24;
    {
        uint64_t s = out_cap ? comp_bytes * 34 / out_cap : 17;
        s = s < FU_SUB_MIN ? FU_SUB_MIN : (s > FU_SUB_MAX ? FU_SUB_MAX : s);
        const uint64_t unit_bytes = ((s - 1) | 1) * 128;
        const uint64_t upc = comp_bytes / K / unit_bytes;                  // units per chunk
        if (upc < 64) warps = 8;      // measured on B200 (tools/dec_shapes.py, 4 bits/symbol): 64 KiB chunks 434 / 528 / 502 GB/s with
        if (upc < 8) warps = 5;       // 24 / 8 / 5 warps, 16 KiB 137 / 204 / 241, 256 KiB 737 / 741 / 737, 1 MiB 933 / 829 / 786
    }
    if (ctx->knobs.fu_warps > 0) warps = (uint32_t)ctx->knobs.fu_warps;   // developer knob
    const int grid_knob = ctx->knobs.fu_grid;   // developer knob
    const unsigned per_sm = warps == 24 ? 1u : (warps == 5 ? 3u : 2u);
    const unsigned grid = grid_knob > 0 ? (unsigned)grid_knob : per_sm * (unsigned)ctx->sm_count;
    void (*kfn)(const FuArgs) = nullptr;
    if (warps == 24) kfn = dec_fused_kernel<24, 1, 1>;
    else if (warps == 8) kfn = dec_fused_kernel<8, 2, 1>;
    else if (warps == 5) kfn = dec_fused_kernel<5, 3, 1>;
    else return hz_fail(ctx, HZ_ERR_ARG, "HZ_FU_WARPS must be 24, 8 or 5");
    // Fewer chunks than CTAs (1 GiB in 16 MiB chunks: 64 chunks for 148 SMs): CTAs share chunks from the start, and a
    // look-back at a unit another CTA holds costs round trips to L2 (look-back 17 % of the warps' time against 9 %).
    // The grid is then launched as CLUSTER PAIRS: the odd CTA joins its peer's chunk and the two push their records into
    // each other's ring through distributed shared memory (developer knob HZ_FU_CLUSTER=0|1 forces either).
    // Measured on B200 (4 bits/symbol, 1 GiB): 64 chunks of 16 MiB 896 -> 968 GB/s, 32 chunks of 32 MiB 868 -> 908, H = 2: 966 -> 1,072;
    // with a chunk or more per CTA the pairs lose (128 chunks of 8 MiB: 1,014 -> 960), so: pairs when every pair can own a chunk.
    bool pairs = warps == 24 && grid_knob <= 0 && 2 * (uint64_t)K <= grid && (grid & 1u) == 0;
    if (ctx->knobs.fu_cluster >= 0) pairs = ctx->knobs.fu_cluster == 1 && warps == 24 && (grid & 1u) == 0;
    if (pairs) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(24 * 32); cfg.dynamicSmemBytes = FU_SMEM_BYTES_CL(24, 2); cfg.stream = ctx->stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        int max_clusters = 0;
        if (cudaOccupancyMaxActiveClusters(&max_clusters, dec_fused_kernel<24, 1, 2>, &cfg) != cudaSuccess) { cudaGetLastError(); max_clusters = 0; }
        if ((unsigned)max_clusters * 2 >= grid) {         // every pair is resident at once (the kernel is persistent)
            hz_prof_begin(ctx);
            cudaError_t e = cudaLaunchKernelEx(&cfg, dec_fused_kernel<24, 1, 2>, a);
            ctx->launches++;
            hz_prof_end(ctx, "dec_fused");
            if (e == cudaSuccess) e = cudaGetLastError();
            if (e != cudaSuccess) return hz_cuda_fail(ctx, e, "dec_fused (cluster pairs)");
        } else {
            pairs = false;
        }
    }
    if (!pairs) HZ_LAUNCH(ctx, "dec_fused", kfn, grid, warps * 32, FU_SMEM_BYTES(warps), a);
#ifdef FU_TIMING
    cudaStreamSynchronize(ctx->stream);
    {
        static const char* nm[12] = {"unit glue", "tma wait", "prepare+lead", "walk", "lookback", "ticket", "output", "skim+glue", "meta+ticket", "table wait", "end barrier", "pick"};
        double tot = 0; for (int i = 0; i < 12; ++i) tot += (double)a.tim[i];
        fprintf(stderr, "FU_TIMING");
        for (int i = 0; i < 12; ++i) fprintf(stderr, "  %s %.1f%%", nm[i], 100.0 * (double)a.tim[i] / tot);
        fprintf(stderr, "\n");
        if (getenv("HZ_FU_TRACE")) {
            static const char* ev[8] = {"chunk", "unit", "staged", "lead", "walk", "lookback", "out", "endbar"};
            const unsigned long long n = a.tim[16] < 4000 ? a.tim[16] : 4000, t0 = n ? a.tim[32] : 0;
            for (unsigned long long i = 0; i < n && i < 600; ++i)
                fprintf(stderr, "TR %8.2f us  w%llu %-8s %llu\n", (double)(a.tim[32 + 2 * i] - t0) / 1965.0, a.tim[33 + 2 * i] >> 48,
                        ev[(a.tim[33 + 2 * i] >> 32) & 7], a.tim[33 + 2 * i] & 0xFFFFFFFFull);
        }
        cudaFree(a.tim);
    }
#endif
    if (dump) {
        cudaStreamSynchronize(ctx->stream);
        if (FILE* f = fopen(dump, "wb")) { fwrite(a.dbg, sizeof(uint32_t), max_units * 32, f); fclose(f); }
        cudaFree(a.dbg);
    }
    *plan_orig_off = P.orig_off; *plan_islice = P.islice;
    return HZ_OK;
}
