// hz_decode.cu — chunked parallel Huffman decode.
//
// Replaces CanonicalHuffman.generateCanonicalCodesFromLengths + TableBasedHuffmanDecoder.decode
// (core/CanonicalHuffman.java:141-146, core/TableBasedHuffmanDecoder.java:36-152, driven by
// CpuCompressionService.decodeChunkParallel, service/cpu/CpuCompressionService.java:511-532).
// A chunk of the .dcz payload is ONE sequential bitstream without restart markers, and the
// container must stay bit-identical, so parallelism inside a chunk comes from the
// self-synchronisation property of Huffman codes:
//
//   plan   (1 CTA)      per-chunk subsequence / sequence / CTA counts and their prefix sums.
//   sync   (many CTAs)  thread i starts HZ_OVERLAP_BITS before subsequence i (a guess), records
//                       where it crosses INTO the subsequence (entry), keeps decoding and
//                       counting symbols until it crosses OUT (exit).  The chain is valid when
//                       exit[i-1] == entry[i] for every i, anchored at bit 0 of the chunk.
//                       Mismatches inside a CTA are repaired by re-decoding from the neighbour's
//                       exit until nothing changes; the CTA's very first subsequence is left to:
//   fix    (1 CTA/chunk) compares every CTA boundary, re-walks from the true position where the
//                       guess was wrong (repeats until stable), then scans symbol counts into
//                       output offsets.
//   write  (many CTAs)  decodes every subsequence again from its verified entry and stores the
//                       symbols; reports HZ_ERR_DECODE if a bit pattern matches no codeword.
//
// Lookup: 2^12-entry shared-memory table (symbol | len<<8) built per CTA from the 256 code
// lengths; longer codes fall back to the canonical first-code walk.  Codes whose used lengths
// are all equal never self-synchronise but need no synchronisation either: entries are computed
// arithmetically.
#include "hz_common.cuh"

#define DT HZ_DEC_THREADS
#define LUTB HZ_DEC_LUT_BITS
#define LUTN (1 << LUTB)

struct __align__(16) DecTables {
    uint16_t lut[LUTN];        // sym | len<<8 ; 0 = not resolvable by the table
    uint32_t first[34];        // first canonical code of each length
    uint32_t count[34];        // symbols per length
    uint32_t offs[34];         // offset of each length in `sorted`
    uint8_t sorted[256];       // symbols ordered by (length, symbol)
    uint8_t len[256];
    int maxlen, minlen, uniform, bad;
};

// Build the decode tables of one chunk; all DT threads participate.
__device__ void build_tables(DecTables& T, const uint8_t* __restrict__ len_k) {
    const uint32_t t = threadIdx.x, lane = t & 31, wid = t >> 5;
    for (uint32_t i = t; i < LUTN / 8; i += DT) reinterpret_cast<uint4*>(T.lut)[i] = make_uint4(0, 0, 0, 0);
    if (t < 34) T.count[t] = 0;
    __syncthreads();
    uint32_t l = t < 256 ? len_k[t] : 0;
    if (t < 256) {
        if (l > 32) l = 33;                       // flagged below
        T.len[t] = (uint8_t)l;
        if (l > 0) atomicAdd(&T.count[l], 1u);
    }
    __syncthreads();
    if (t == 0) {
        uint32_t c = 0, o = 0;
        int mx = 0, mn = 99;
        uint64_t kraft = 0;                       // in units of 2^-32
        T.first[0] = 0; T.offs[0] = 0;
        for (int L = 1; L <= 32; ++L) {
            c = (c + (L > 1 ? T.count[L - 1] : 0u)) << 1;
            T.first[L] = c; T.offs[L] = o; o += T.count[L];
            if (T.count[L]) { mx = L; if (mn == 99) mn = L; kraft += (uint64_t)T.count[L] << (32 - L); }
        }
        T.maxlen = mx; T.minlen = mn == 99 ? 0 : mn;
        T.uniform = (mx > 0 && mx == mn) ? mx : 0;
        T.bad = (T.count[33] != 0) || (kraft > (1ull << 32));
    }
    __syncthreads();
    if (t < 256 && l > 0 && l <= 32) {
        uint32_t rank = 0;
        for (uint32_t s = 0; s < t; ++s) rank += (T.len[s] == l);
        T.sorted[T.offs[l] + rank] = (uint8_t)t;
    }
    __syncthreads();
    if (T.bad) return;
    // warp-cooperative table fill: warp w takes symbols w, w+8, ...
    for (uint32_t s = wid; s < 256; s += DT / 32) {
        uint32_t ls = T.len[s];
        if (ls == 0 || ls > LUTB) continue;
        uint32_t rank = 0;   // recompute code = first + rank via sorted position
        // position of s inside its length class
        // (sorted[] is ordered, so binary search is possible; a linear warp vote is simpler)
        uint32_t cnt = T.count[ls], base_o = T.offs[ls];
        for (uint32_t j = lane; j < cnt; j += 32) if (T.sorted[base_o + j] == s) rank = j + 1;
        rank = __reduce_max_sync(0xffffffffu, rank) - 1;
        uint32_t code = T.first[ls] + rank;
        uint32_t span = 1u << (LUTB - ls), b = code << (LUTB - ls);
        uint16_t e = (uint16_t)(s | (ls << 8));
        for (uint32_t x = lane; x < span; x += 32) T.lut[b + x] = e;
    }
    __syncthreads();
}

// Bit reader over one chunk: 64-bit window, zero bits past the end of the chunk.
struct BitWin {
    const uint8_t* base;     // first byte of the chunk
    uint32_t csize;          // bytes in the chunk
    uint64_t win;            // upcoming bits, MSB first
    int avail;               // valid bits in win
    uint64_t next;           // next byte offset to fetch (multiple of 4 relative to `algn`)
    int algn;                // base address & 3
};

__device__ __forceinline__ uint32_t fetch_word(const BitWin& r, int64_t boff) {
    // 4 bytes at chunk byte offset boff (may be negative / past the end -> zeros), big-endian
    if (boff >= 0 && boff + 4 <= (int64_t)r.csize) {
        uint32_t w = *reinterpret_cast<const uint32_t*>(r.base + boff);   // aligned by construction
        return bswap32(w);
    }
    uint32_t w = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int64_t b = boff + i;
        uint32_t v = (b >= 0 && b < (int64_t)r.csize) ? r.base[b] : 0;
        w = (w << 8) | v;
    }
    return w;
}

__device__ __forceinline__ void bw_seek(BitWin& r, uint64_t bitpos) {
    // word-aligned (in address space) fetches: aligned byte offset = ((algn + bitpos/8) & ~3) - algn
    int64_t byte = (int64_t)(bitpos >> 3);
    int64_t a = ((byte + r.algn) & ~(int64_t)3) - r.algn;
    uint32_t skip = (uint32_t)((byte - a) * 8 + (bitpos & 7));       // 0..31
    uint64_t w0 = fetch_word(r, a), w1 = fetch_word(r, a + 4);
    r.win = ((w0 << 32) | w1) << skip;
    r.avail = 64 - (int)skip;
    r.next = (uint64_t)(a + 8);
}

__device__ __forceinline__ void bw_refill(BitWin& r) {
    if (r.avail <= 32) {
        uint64_t w = fetch_word(r, (int64_t)r.next);
        r.win |= w << (32 - r.avail);
        r.avail += 32;
        r.next += 4;
    }
}

// Decode one codeword from the window.  Returns its length (>=1) and the symbol;
// an unmatched pattern consumes 1 bit and returns sym = -1.
__device__ __forceinline__ int decode_one(const DecTables& T, uint64_t win, int* sym) {
    uint32_t e = T.lut[(uint32_t)(win >> (64 - LUTB))];
    if (e) { *sym = e & 0xFF; return e >> 8; }
    const uint32_t top = (uint32_t)(win >> 32);
    for (int l = LUTB + 1; l <= T.maxlen; ++l) {
        uint32_t c = top >> (32 - l);
        uint32_t d = c - T.first[l];
        if (c >= T.first[l] && d < T.count[l]) { *sym = T.sorted[T.offs[l] + d]; return l; }
    }
    *sym = -1;
    return 1;
}

struct ChunkGeom {
    uint64_t comp_off; uint32_t comp_size; uint32_t orig_size; uint64_t orig_off;
    uint32_t nsub, nseq, sub_base, seq_base, cta_base, ncta;
};

// plan arrays (SoA, K+1 entries each where a total is needed)
struct DecPlan {
    uint32_t* nsub; uint32_t* sub_base; uint32_t* seq_base; uint32_t* cta_base; uint64_t* orig_off;
};

__device__ __forceinline__ uint32_t ceil_div_u64(uint64_t a, uint32_t b) { return (uint32_t)((a + b - 1) / b); }

__global__ void __launch_bounds__(1024)
dec_plan_kernel(const uint32_t* __restrict__ comp_size, const uint32_t* __restrict__ orig_size,
                const uint64_t* __restrict__ orig_off_in, uint32_t K, DecPlan P) {
    __shared__ uint64_t part[4][1024];
    const uint32_t t = threadIdx.x;
    const uint32_t per = (K + 1023) / 1024;
    const uint32_t lo = min(K, t * per), hi = min(K, lo + per);
    uint64_t s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    for (uint32_t i = lo; i < hi; ++i) {
        uint32_t ns = orig_size[i] ? max(1u, ceil_div_u64((uint64_t)comp_size[i] * 8, HZ_SUB_BITS)) : 0;
        uint32_t nq = (ns + DT - 1) / DT;
        s0 += ns; s1 += nq; s2 += (nq + HZ_SEQ_PER_CTA - 1) / HZ_SEQ_PER_CTA; s3 += orig_size[i];
    }
    part[0][t] = s0; part[1][t] = s1; part[2][t] = s2; part[3][t] = s3;
    __syncthreads();
    if (t < 4) {
        uint64_t a = 0;
        for (int j = 0; j < 1024; ++j) { uint64_t x = part[t][j]; part[t][j] = a; a += x; }
        if (t == 0) P.sub_base[K] = (uint32_t)a;
        if (t == 1) P.seq_base[K] = (uint32_t)a;
        if (t == 2) P.cta_base[K] = (uint32_t)a;
        if (t == 3) P.orig_off[K] = a;
    }
    __syncthreads();
    s0 = part[0][t]; s1 = part[1][t]; s2 = part[2][t]; s3 = part[3][t];
    for (uint32_t i = lo; i < hi; ++i) {
        uint32_t ns = orig_size[i] ? max(1u, ceil_div_u64((uint64_t)comp_size[i] * 8, HZ_SUB_BITS)) : 0;
        uint32_t nq = (ns + DT - 1) / DT;
        P.nsub[i] = ns; P.sub_base[i] = (uint32_t)s0; P.seq_base[i] = (uint32_t)s1; P.cta_base[i] = (uint32_t)s2;
        P.orig_off[i] = orig_off_in ? orig_off_in[i] : s3;
        s0 += ns; s1 += nq; s2 += (nq + HZ_SEQ_PER_CTA - 1) / HZ_SEQ_PER_CTA; s3 += orig_size[i];
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

// Decode subsequence `i` of a chunk starting at absolute bit `start` (>= nominal unless run-in):
// counts the codewords that BEGIN in [max(start, nominal_i), nominal_{i+1}) and returns the exit
// offset (first codeword boundary at or after nominal_{i+1}, relative to it).  If `entry` is
// non-null the run-in phase is performed first and *entry receives the crossing offset.
__device__ __forceinline__ void scan_subseq(const DecTables& T, BitWin& r, uint64_t start, uint64_t nominal,
                                            uint32_t* entry, uint32_t* count, uint32_t* exitv) {
    uint64_t pos = start;
    bw_seek(r, pos);
    int sym;
    if (entry) {
        while (pos < nominal) {
            bw_refill(r);
            int l = decode_one(T, r.win, &sym);
            r.win <<= l; r.avail -= l; pos += l;
        }
        *entry = (uint32_t)(pos - nominal);
    }
    const uint64_t end = nominal + HZ_SUB_BITS;
    uint32_t cnt = 0;
    while (pos < end) {
        bw_refill(r);
        int l = decode_one(T, r.win, &sym);
        r.win <<= l; r.avail -= l; pos += l;
        ++cnt;
    }
    *count = cnt;
    *exitv = (uint32_t)(pos - end);
}

// record layout: entry (8) | exit (8) | count (16)
__device__ __forceinline__ uint32_t pack_rec(uint32_t entry, uint32_t exitv, uint32_t count) {
    return entry | (exitv << 8) | (count << 16);
}

__global__ void __launch_bounds__(DT)
dec_sync_kernel(const uint8_t* __restrict__ comp, const uint64_t* __restrict__ comp_off,
                const uint32_t* __restrict__ comp_size, const uint8_t* __restrict__ len_tab,
                uint32_t K, DecPlan P, uint32_t* __restrict__ rec, uint32_t* __restrict__ seqcnt, int* status) {
    __shared__ DecTables T;
    __shared__ uint32_t s_exit[DT];
    __shared__ uint32_t s_red[DT / 32];
    __shared__ uint32_t s_k;
    const uint32_t t = threadIdx.x, lane = t & 31, wid = t >> 5;
    if (blockIdx.x >= P.cta_base[K]) return;
    if (t == 0) s_k = find_chunk(P.cta_base, K, blockIdx.x);
    __syncthreads();
    const uint32_t k = s_k;
    build_tables(T, len_tab + (size_t)k * 256);
    if (T.bad) { if (t == 0) hz_set_status(status, HZ_ERR_BAD_LENGTHS); return; }

    BitWin r;
    r.base = comp + comp_off[k];
    r.csize = comp_size[k];
    r.algn = (int)(reinterpret_cast<uintptr_t>(r.base) & 3);
    const uint32_t nsub = P.nsub[k];
    const uint32_t nseq = (nsub + DT - 1) / DT;
    const uint32_t cta_in_chunk = blockIdx.x - P.cta_base[k];
    const uint32_t U = (uint32_t)T.uniform;

    uint32_t carry_exit = 0;         // exit of the previous sequence's last subsequence (q > 0)
    for (uint32_t q = 0; q < HZ_SEQ_PER_CTA; ++q) {
        const uint32_t sq = cta_in_chunk * HZ_SEQ_PER_CTA + q;
        if (sq >= nseq) break;
        const uint32_t i = sq * DT + t;
        const bool active = i < nsub;
        const uint64_t nominal = (uint64_t)i * HZ_SUB_BITS;
        uint32_t entry = 0, count = 0, exitv = 0;
        if (active) {
            if (U) {
                // equal-length code: boundaries are the multiples of U
                entry = (uint32_t)((U - nominal % U) % U);
                scan_subseq(T, r, nominal + entry, nominal, nullptr, &count, &exitv);
            } else if (i == 0) {
                scan_subseq(T, r, 0, 0, nullptr, &count, &exitv);
            } else if (t == 0 && q > 0) {
                entry = carry_exit;
                scan_subseq(T, r, nominal + entry, nominal, nullptr, &count, &exitv);
            } else {
                uint64_t start = nominal > HZ_OVERLAP_BITS ? nominal - HZ_OVERLAP_BITS : 0;
                scan_subseq(T, r, start, nominal, &entry, &count, &exitv);
            }
        }
        s_exit[t] = exitv;
        __syncthreads();
        // repair mismatches inside the sequence until the chain is consistent
        if (!U) {
            for (;;) {
                bool fix = false;
                uint32_t want = 0;
                if (active && t > 0) { want = s_exit[t - 1]; fix = want != entry; }
                if (!__syncthreads_or(fix)) break;
                if (fix) {
                    entry = want;
                    scan_subseq(T, r, nominal + entry, nominal, nullptr, &count, &exitv);
                    s_exit[t] = exitv;
                }
                __syncthreads();
            }
        }
        if (active) rec[P.sub_base[k] + i] = pack_rec(entry, exitv, count);
        // symbols of this sequence
        uint32_t c = active ? count : 0;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
        if (lane == 0) s_red[wid] = c;
        __syncthreads();
        if (t == 0) {
            uint32_t a = 0;
            for (int w = 0; w < DT / 32; ++w) a += s_red[w];
            seqcnt[P.seq_base[k] + sq] = a;
        }
        carry_exit = s_exit[DT - 1];
        __syncthreads();
    }
}

// One CTA per chunk: repair CTA boundaries, then turn per-sequence symbol counts into offsets.
__global__ void __launch_bounds__(DT)
dec_fix_kernel(const uint8_t* __restrict__ comp, const uint64_t* __restrict__ comp_off,
               const uint32_t* __restrict__ comp_size, const uint8_t* __restrict__ len_tab,
               DecPlan P, uint32_t* __restrict__ rec, uint32_t* __restrict__ seqcnt, int* status) {
    __shared__ DecTables T;
    __shared__ uint32_t s_warp[DT / 32 + 1];
    const uint32_t k = blockIdx.x, t = threadIdx.x, lane = t & 31, wid = t >> 5;
    const uint32_t nsub = P.nsub[k];
    if (nsub == 0) return;
    const uint32_t nseq = (nsub + DT - 1) / DT;
    const uint32_t ncta = (nseq + HZ_SEQ_PER_CTA - 1) / HZ_SEQ_PER_CTA;
    uint32_t* R = rec + P.sub_base[k];
    uint32_t* SC = seqcnt + P.seq_base[k];
    const uint32_t SUBS_PER_CTA = DT * HZ_SEQ_PER_CTA;

    if (ncta > 1) {
        // quick check first: any boundary mismatch at all?
        bool any = false;
        for (uint32_t b = 1 + t; b < ncta; b += DT) {
            uint32_t i = b * SUBS_PER_CTA;
            any |= ((R[i - 1] >> 8) & 0xFF) != (R[i] & 0xFF);
        }
        if (__syncthreads_or(any)) {
            build_tables(T, len_tab + (size_t)k * 256);
            if (T.bad) return;
            BitWin r;
            r.base = comp + comp_off[k];
            r.csize = comp_size[k];
            r.algn = (int)(reinterpret_cast<uintptr_t>(r.base) & 3);
            for (;;) {
                bool changed = false;
                for (uint32_t b0 = 1; b0 < ncta; b0 += DT) {
                    const uint32_t b = b0 + t;
                    uint32_t want = 0, i = 0;
                    bool walk = false;
                    if (b < ncta) {
                        i = b * SUBS_PER_CTA;
                        want = (R[i - 1] >> 8) & 0xFF;
                        walk = want != (R[i] & 0xFF);
                    }
                    __syncthreads();           // all reads of neighbours' exits before any update
                    if (walk) {
                        const uint32_t iend = min(nsub, i + SUBS_PER_CTA);
                        uint32_t entry = want;
                        for (; i < iend; ++i) {
                            const uint64_t nominal = (uint64_t)i * HZ_SUB_BITS;
                            uint32_t count, exitv;
                            scan_subseq(T, r, nominal + entry, nominal, nullptr, &count, &exitv);
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

__global__ void __launch_bounds__(DT)
dec_write_kernel(const uint8_t* __restrict__ comp, const uint64_t* __restrict__ comp_off,
                 const uint32_t* __restrict__ comp_size, const uint32_t* __restrict__ orig_size,
                 const uint8_t* __restrict__ len_tab, uint32_t K, DecPlan P,
                 const uint32_t* __restrict__ rec, const uint32_t* __restrict__ seqoff,
                 uint8_t* __restrict__ out, uint64_t out_cap, int* status) {
    __shared__ DecTables T;
    __shared__ uint32_t s_warp[DT / 32 + 1];
    __shared__ uint32_t s_k;
    const uint32_t t = threadIdx.x, lane = t & 31, wid = t >> 5;
    if (blockIdx.x >= P.cta_base[K]) return;
    if (t == 0) s_k = find_chunk(P.cta_base, K, blockIdx.x);
    __syncthreads();
    const uint32_t k = s_k;
    const uint32_t osize = orig_size[k];
    const uint64_t ooff = P.orig_off[k];
    if (ooff + osize > out_cap) { if (t == 0) hz_set_status(status, HZ_ERR_OUT_TOO_SMALL); return; }
    build_tables(T, len_tab + (size_t)k * 256);
    if (T.bad) { if (t == 0) hz_set_status(status, HZ_ERR_BAD_LENGTHS); return; }
    BitWin r;
    r.base = comp + comp_off[k];
    r.csize = comp_size[k];
    r.algn = (int)(reinterpret_cast<uintptr_t>(r.base) & 3);
    const uint32_t nsub = P.nsub[k];
    const uint32_t nseq = (nsub + DT - 1) / DT;
    const uint32_t cta_in_chunk = blockIdx.x - P.cta_base[k];
    uint8_t* o = out + ooff;

    for (uint32_t q = 0; q < HZ_SEQ_PER_CTA; ++q) {
        const uint32_t sq = cta_in_chunk * HZ_SEQ_PER_CTA + q;
        if (sq >= nseq) break;
        const uint32_t i = sq * DT + t;
        const bool active = i < nsub;
        uint32_t rv = active ? rec[P.sub_base[k] + i] : 0;
        uint32_t count = rv >> 16;
        // exclusive scan of counts inside the sequence
        uint32_t inc = count;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t x = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += x;
        }
        if (lane == 31) s_warp[wid] = inc;
        __syncthreads();
        if (t == 0) {
            uint32_t a = 0;
            for (int w = 0; w < DT / 32; ++w) { uint32_t x = s_warp[w]; s_warp[w] = a; a += x; }
        }
        __syncthreads();
        uint32_t obase = seqoff[P.seq_base[k] + sq] + s_warp[wid] + inc - count;
        __syncthreads();
        if (!active) continue;
        // the last subsequence of the chunk runs until orig_size symbols exist (bits past the end
        // of the chunk read as zero, TableBasedHuffmanDecoder.java:204-208)
        uint32_t todo = count;
        if (i == nsub - 1) todo = osize > obase ? osize - obase : 0;
        else if (obase >= osize) todo = 0;
        else if (obase + todo > osize) todo = osize - obase;
        if (todo == 0) continue;
        uint64_t pos = (uint64_t)i * HZ_SUB_BITS + (rv & 0xFF);
        bw_seek(r, pos);
        uint8_t* dst = o + obase;
        bool err = false;
        for (uint32_t j = 0; j < todo; ++j) {
            bw_refill(r);
            int sym;
            int l = decode_one(T, r.win, &sym);
            r.win <<= l; r.avail -= l;
            if (sym < 0) { err = true; sym = 0; }
            dst[j] = (uint8_t)sym;
        }
        if (err) hz_set_status(status, HZ_ERR_DECODE);
    }
}

int hzk_decode(hz_ctx* ctx, const uint8_t* d_comp, uint64_t comp_bytes, const uint64_t* d_comp_off,
               const uint32_t* d_comp_size, const uint32_t* d_orig_size, const uint64_t* d_orig_off,
               const uint8_t* d_len, uint32_t K, uint8_t* d_out, uint64_t out_cap) {
    if (K == 0) return HZ_OK;
    // plan arrays
    HZ_TRY(hz_reserve(ctx, &ctx->dec_meta, ((size_t)K + 1) * (4 * sizeof(uint32_t) + sizeof(uint64_t)) + 64));
    DecPlan P;
    uint8_t* m = (uint8_t*)ctx->dec_meta.p;
    P.orig_off = (uint64_t*)m; m += ((size_t)K + 1) * sizeof(uint64_t);
    P.nsub = (uint32_t*)m; m += ((size_t)K + 1) * sizeof(uint32_t);
    P.sub_base = (uint32_t*)m; m += ((size_t)K + 1) * sizeof(uint32_t);
    P.seq_base = (uint32_t*)m; m += ((size_t)K + 1) * sizeof(uint32_t);
    P.cta_base = (uint32_t*)m;
    HZ_LAUNCH(ctx, "dec_plan", dec_plan_kernel, 1, 1024, 0, d_comp_size, d_orig_size, d_orig_off, K, P);
    // upper bounds (no host sync): every chunk has at most ceil(comp_size*8/SUB_BITS)+1 subsequences
    const uint64_t max_sub = comp_bytes * 8 / HZ_SUB_BITS + 2ull * K + 2;
    const uint64_t max_seq = max_sub / DT + K + 1;
    const uint64_t max_cta = max_seq / HZ_SEQ_PER_CTA + K + 1;
    if (max_cta > 0x7fffffffull) return hz_fail(ctx, HZ_ERR_ARG, "decode grid too large");
    HZ_TRY(hz_reserve(ctx, &ctx->dec_rec, max_sub * sizeof(uint32_t)));
    HZ_TRY(hz_reserve(ctx, &ctx->dec_seqcnt, max_seq * sizeof(uint32_t)));
    uint32_t* rec = (uint32_t*)ctx->dec_rec.p;
    uint32_t* seqcnt = (uint32_t*)ctx->dec_seqcnt.p;
    HZ_LAUNCH(ctx, "dec_sync", dec_sync_kernel, (unsigned)max_cta, DT, 0,
              d_comp, d_comp_off, d_comp_size, d_len, K, P, rec, seqcnt, ctx->d_status);
    HZ_LAUNCH(ctx, "dec_fix", dec_fix_kernel, K, DT, 0,
              d_comp, d_comp_off, d_comp_size, d_len, P, rec, seqcnt, ctx->d_status);
    HZ_LAUNCH(ctx, "dec_write", dec_write_kernel, (unsigned)max_cta, DT, 0,
              d_comp, d_comp_off, d_comp_size, d_orig_size, d_len, K, P, rec, seqcnt, d_out, out_cap, ctx->d_status);
    return HZ_OK;
}
