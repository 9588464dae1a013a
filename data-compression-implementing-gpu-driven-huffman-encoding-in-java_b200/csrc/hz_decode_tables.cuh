// hz_decode_tables.cuh — lookup tables and shared-memory helpers shared by the decode kernels
// (hz_decode.cu: identity / legacy multi-pass kernels; hz_decode_fused.cu: single-residency decoder).
#pragma once
#include <cstdlib>
#include "hz_common.cuh"

#define DT 256
#define LUTB 12
#define LUTN (1 << LUTB)
#define DEC_NO_TABLE 0xFFFFFFFFu
#define DEC_TAB_W 0                              // uint2 wlut[LUTN]
#define DEC_TAB_S (LUTN * 8)                    // uint16 slut[LUTN]
#define DEC_TAB_AUX (LUTN * 8 + LUTN * 2)       // DecAux
#define DEC_TABLE_BYTES (DEC_TAB_AUX + 1024)

struct __align__(16) DecAux {
    uint64_t lim[34];          // exclusive upper bound of the left-justified (32-bit) codes of each length
    int32_t symbase[34];       // sorted[symbase[l] + code] = symbol of a length-l code
    uint8_t sorted[256];       // symbols ordered by (length, symbol)
    uint8_t len[256];          // code length of every symbol
    int maxlen, minlen, uniform, bad;
};
static_assert(sizeof(DecAux) <= 1024, "DecAux must fit its 1 KiB slot");
// shared-memory offset of DecAux::sorted relative to the write kernel's wlut (WriteShared: wlut, then aux)
#define DEC_W_SORTED_REL (LUTN * 8 + (uint32_t)offsetof(DecAux, sorted))

// ---------------------------------------------------------------------------------------------
// table construction (all DT threads).  scratch: >= 8 KiB + 2 KiB of shared memory.
// ---------------------------------------------------------------------------------------------
#define DEC_BUILD_SCRATCH (LUTN * 2 + 512 + 256 + 8 * 34 * 4 + 3 * 34 * 4)

// length of the (long) code that starts the left-justified 32 stream bits v, searched in
// [lmin, lmax]; 0 = no code matches
__device__ __forceinline__ uint32_t long_len(const DecAux& A, uint32_t v, uint32_t lmin, uint32_t lmax) {
    if (lmin == 0) return 0;
    uint32_t l = lmin;
    while (l < lmax && (uint64_t)v >= A.lim[l]) ++l;
    return (uint64_t)v < A.lim[l] ? l : 0;
}

// barrier of the DT threads that build a table (threads 0..DT-1 of the CTA; the write kernel's CTA is larger)
__device__ __forceinline__ void bt_sync() { asm volatile("bar.sync 1, %0;" ::"n"(DT) : "memory"); }

// FMT 0: wlut of the legacy write kernel (.y = ltot | 8n << 16).  FMT 1: wlut of the fused kernel, whose packed
// counter keeps the stream position in the HIGH half: .y = ltot << 16 | 8n; a single-length long code (13..24 bits)
// has .y = l << 16 | 8 - so .y >= (LUTB + 1) << 16 is the predicate of the second-level load - and
// .x = (DEC_W_SORTED_REL + symbase[l]) << 5 | (32 - l): the low five bits are the right shift that leaves the first l
// stream bits, the rest (arithmetic shift) the wlut-relative address of sorted[symbase[l]]; entries with several
// candidate lengths / no code have .y = 0x40000000 (the walk's end compare catches it).
template <bool WANT_W, bool WANT_S, int FMT = 0>
__device__ void build_tables(DecAux& A, uint2* __restrict__ wlut, uint16_t* __restrict__ slut,
                             uint8_t* __restrict__ scratch, const uint8_t* __restrict__ len_k) {
    uint16_t* base = reinterpret_cast<uint16_t*>(scratch);                // [LUTN] sym | len<<8
    uint32_t* cntw = reinterpret_cast<uint32_t*>(scratch + LUTN * 2 + 768);   // [8][34]
    uint32_t* first = cntw + 8 * 34;                                      // [34] first canonical code per length
    uint32_t* count = first + 34;                                         // [34] symbols per length
    uint32_t* offs = count + 34;                                          // [34] offset of each length in sorted[]
    const uint32_t t = threadIdx.x, lane = t & 31, wid = t >> 5;
    uint32_t l = len_k[t];
    if (l > 32) l = 33;
    A.len[t] = (uint8_t)l;
    for (uint32_t i = t; i < 8 * 34; i += DT) cntw[i] = 0;
    bt_sync();
    const uint32_t same = __match_any_sync(0xffffffffu, l);
    const uint32_t rank_w = __popc(same & ((1u << lane) - 1));
    if (rank_w == 0) cntw[wid * 34 + l] = __popc(same);
    bt_sync();
    if (t < 34) {
        uint32_t c = 0;
        for (int w = 0; w < 8; ++w) c += cntw[w * 34 + t];
        count[t] = t == 0 ? 0 : c;
    }
    bt_sync();
    if (t == 0) {
        uint32_t c = 0, o = 0;
        int mx = 0, mn = 0;
        uint64_t kraft = 0;                       // in units of 2^-32
        first[0] = 0; offs[0] = 0; A.lim[0] = 0; A.symbase[0] = 0;
        for (int L = 1; L <= 32; ++L) {
            c = (c + (L > 1 ? count[L - 1] : 0u)) << 1;
            first[L] = c; offs[L] = o; o += count[L];
            A.symbase[L] = (int32_t)offs[L] - (int32_t)c;
            A.lim[L] = ((uint64_t)c + count[L]) << (32 - L);
            if (count[L]) { mx = L; if (!mn) mn = L; kraft += (uint64_t)count[L] << (32 - L); }
        }
        offs[33] = o; first[33] = 0; A.lim[33] = 0; A.symbase[33] = 0;
        A.maxlen = mx; A.minlen = mn;
        A.uniform = (mx > 0 && mx == mn) ? mx : 0;
        A.bad = (count[33] != 0) || (kraft > (1ull << 32));
    }
    bt_sync();
    if (A.bad) return;
    if (l >= 1 && l <= 32) {
        uint32_t rank = rank_w;
        for (uint32_t w = 0; w < wid; ++w) rank += cntw[w * 34 + l];
        A.sorted[offs[l] + rank] = (uint8_t)t;
    }
    bt_sync();
    // single-symbol table.  The left-justified codes of a canonical code are ordered by length, so the code
    // that starts a 12-bit prefix x has the first length l with x < lim12[l] (lim12 = A.lim >> 20, exact for
    // l <= LUTB): LUTB register compares per entry, no search.  Thread t fills entries t, t + 256, ...
    // (consecutive lanes -> consecutive entries: conflict-free shared accesses, coalesced table stores).
    {
        uint32_t lim12[LUTB + 1];
#pragma unroll
        for (int L = 1; L <= LUTB; ++L) lim12[L] = (uint32_t)(A.lim[L] >> (32 - LUTB));
#pragma unroll 4
        for (uint32_t x = t; x < LUTN; x += DT) {
            uint32_t li = 1;
#pragma unroll
            for (int L = 1; L <= LUTB; ++L) li += x >= lim12[L];
            uint16_t e = 0;
            if (li <= LUTB) e = (uint16_t)(A.sorted[A.symbase[li] + (int32_t)(x >> (LUTB - li))] | (li << 8));
            base[x] = e;
        }
    }
    bt_sync();
    // multi-symbol tables
    const uint32_t maxlen = (uint32_t)A.maxlen;
    for (uint32_t x = t; x < LUTN; x += DT) {
        const uint32_t e0 = base[x];
        // FMT 1: the fused kernel's predicated long-code load also runs (harmlessly) on entries with several
        // candidate lengths / no code, so their .x keeps 31 in the low five bits (address offset <= 1 + lmax << 6)
        uint2 we = make_uint2(FMT ? 31u : 0u, FMT ? 0x40000000u : 0xC0000000u);
        uint32_t se = 0;
        if (e0) {
            const uint32_t l0 = e0 >> 8;
            uint32_t syms = e0 & 0xFF, used = l0, n = 1, wtot = l0, wn = 1, cur = x, lprev = l0;
            for (;;) {
                cur = (cur << lprev) & (LUTN - 1);
                const uint32_t e = base[cur];
                if (!e) break;
                const uint32_t le = e >> 8;
                if (used + le > LUTB) break;
                if (n < 4) { syms |= (e & 0xFF) << (8 * n); wtot = used + le; wn = n + 1; }
                used += le; ++n; lprev = le;
                if (!WANT_S && n == 4) break;             // (the count table alone looks beyond four symbols)
            }
            we = FMT ? make_uint2(syms, (wtot << 16) | (wn << 3)) : make_uint2(syms, wtot | (wn << 19));
            se = used | (l0 << 6) | (n << 12);
        } else if (maxlen > LUTB) {
            // the prefix starts a code longer than LUTB bits (or no code): candidate lengths at both ends
            const uint32_t vlo = x << (32 - LUTB), vhi = vlo | ((1u << (32 - LUTB)) - 1);
            const uint32_t lmin = long_len(A, vlo, LUTB + 1, maxlen);
            if (lmin) {
                uint32_t lmax = long_len(A, vhi, lmin, maxlen);
                if (!lmax) lmax = maxlen;
                we.x = FMT ? (31u | (lmin << 5) | (lmax << 11)) : (lmin | (lmax << 6));
                // every code under this prefix has the same length: `sorted` is the second-level table
                if (FMT) {
                    if (lmin == lmax && lmin <= 24)
                        we = make_uint2(((DEC_W_SORTED_REL + (uint32_t)A.symbase[lmin]) << 5) | (32 - lmin), (lmin << 16) | 8u);
                } else if (lmin == lmax && lmin < 32)
                    we = make_uint2(DEC_W_SORTED_REL + (uint32_t)A.symbase[lmin], 0x80000000u | lmin | (8u << 16));
                se = lmin == lmax ? (lmin | (lmin << 6) | (1u << 12)) : (lmin | (lmax << 6));
            }
        }
        if (WANT_W) wlut[x] = we;
        if (WANT_S) slut[x] = (uint16_t)se;
    }
    bt_sync();
}

// copy `bytes` (multiple of 16) from global to shared with all threads of the CTA
__device__ __forceinline__ void copy_g2s16(void* dst, const void* src, uint32_t bytes) {
    const uint4* s = reinterpret_cast<const uint4*>(src);
    uint4* d = reinterpret_cast<uint4*>(dst);
    for (uint32_t i = threadIdx.x; i < bytes / 16; i += blockDim.x) d[i] = s[i];
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

// shared-state-space accesses with 32-bit addresses (keeps the hot loops free of generic->shared
// window arithmetic)
__device__ __forceinline__ uint32_t lds32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint2 lds64(uint32_t a) { uint2 v; asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a)); return v; }
__device__ __forceinline__ uint32_t lds16(uint32_t a) { uint16_t v; asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint32_t lds8(uint32_t a) { uint32_t v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ void sts32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts8(uint32_t a, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }

// keep a value in a register: stops the compiler from re-deriving shared-window bases in hot loops
__device__ __forceinline__ uint32_t pin_reg(uint32_t v) { asm volatile("" : "+r"(v)); return v; }

// PTX twin of long_len() for the hand-written loops.  Registers of the enclosing asm block: v (32 stream
// bits), l (in: shortest, out: length found or 0), m (longest candidate), auxb (shared address of DecAux);
// temporaries a, t, u and predicate pq.  SFX makes the labels unique.
#define HZ_PTX_LONGLEN(SFX)                         \
    "setp.eq.u32 pq, l, 0;\n"                       \
    "@pq bra HZL_END" SFX ";\n"                     \
    "mad.lo.u32 a, l, 8, auxb;\n"                   \
    "HZL_TOP" SFX ":\n"                             \
    "ld.shared.v2.u32 {t, u}, [a];\n"               \
    "setp.ne.u32 pq, u, 0;\n"                       \
    "@pq bra HZL_END" SFX ";\n"                     \
    "setp.lt.u32 pq, v, t;\n"                       \
    "@pq bra HZL_END" SFX ";\n"                     \
    "setp.ge.u32 pq, l, m;\n"                       \
    "@pq mov.u32 l, 0;\n"                           \
    "@pq bra HZL_END" SFX ";\n"                     \
    "add.u32 l, l, 1;\n"                            \
    "add.u32 a, a, 8;\n"                            \
    "bra HZL_TOP" SFX ";\n"                         \
    "HZL_END" SFX ":\n"
static_assert(offsetof(DecAux, lim) == 0, "HZ_PTX_LONGLEN reads lim[] at the start of DecAux");
