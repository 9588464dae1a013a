// hz_codebook.cu — canonical Huffman codebooks, one CTA per chunk (stage 2 of encode).
//
// Replaces CanonicalHuffman.buildCanonicalCodes(long[256]) (core/CanonicalHuffman.java:19-132).
// The code LENGTHS of the reference depend on how java.util.PriorityQueue orders
// equal-frequency internal nodes (HuffmanNode.compareTo, core/HuffmanNode.java:52-58, returns 0
// for two internal nodes of equal frequency), so the tree is built by replaying the JDK's binary heap
// operation by operation: offer = siftUp (stop when cmp(x,parent) >= 0), poll = move last to root +
// siftDown (pick right child only if cmp(left,right) > 0; stop when cmp(x,child) <= 0) - by one warp per
// chunk (warp_heap_replay) or, for thousands of small chunks, one lane per chunk (codebook_lane_kernel).
// A heap entry is one uint64: (freq << 18) | ((symbol+1) << 9) | node_id — comparing
// (entry >> 9) is exactly compareTo (internal nodes carry symbol -1 -> field 0).
// Everything around the serial heap (summing segment histograms, leaf depths by parent chasing,
// canonical code assignment, exact compressed sizes and per-segment output bit offsets) is
// parallel over the CTA's 256 threads.
#include <cstdlib>
#include <cstring>
#include "hz_common.cuh"
#include "hz_hist_lanes.cuh"

#define CB_THREADS 256

// compareTo on heap entries: key(a) > key(b) with key = entry >> 9  <=>  a > (b | 511) (the low 9 bits, the node
// id, are saturated away) - one 64-bit compare instead of two 64-bit shifts and a compare
__device__ __forceinline__ bool key_gt(uint64_t a, uint64_t b) { return a > (b | 511ull); }

__device__ __forceinline__ void heap_offer(uint64_t* q, int& size, uint64_t x) {
    int k = size++;
    const uint64_t xs = x | 511ull;
    while (k > 0) {
        int parent = (k - 1) >> 1;
        uint64_t e = q[parent];
        if (!(e > xs)) break;                       // cmp(x, parent) >= 0
        q[k] = e;
        k = parent;
    }
    q[k] = x;
}

__device__ __forceinline__ uint64_t heap_poll(uint64_t* q, int& size) {
    uint64_t result = q[0];
    int n = --size;
    if (n > 0) {
        uint64_t x = q[n];
        int k = 0;
        const int half = n >> 1;
        while (k < half) {
            int child = 2 * k + 1;
            uint64_t c = q[child];
            int right = child + 1;
            if (right < n) {
                uint64_t r = q[right];
                if (key_gt(c, r)) { c = r; child = right; }     // cmp(left, right) > 0
            }
            if (!key_gt(x, c)) break;                           // cmp(x, child) <= 0
            q[k] = c;
            k = child;
        }
        q[k] = x;
    }
    return result;
}

// ---------------------------------------------------------------------------------------------
// The same heap replayed by a WARP (HZ_CB_PLAIN keeps the one-thread loops above for A/B runs).
// A literal siftDown is 7 dependent rounds of load -> compare -> select -> compare -> store.  But the
// children it visits - the heap's min-child path - do not depend on the element x being sifted:
//   1. every lane compares the two children of 4 inner slots (one 128-bit load each); four ballots give all
//      lanes the 127 "right child is the smaller one" bits (cmp(left, right) > 0, as PriorityQueue.siftDown);
//   2. the path follows from the bits alone (register arithmetic, no memory access);
//   3. lane L loads the path entry of level L + 1 and compares it with x; a ballot gives the number t of
//      levels x sinks, lanes < t move their entry up one level, lane t stores x.
// siftUp is one round as well: lane i loads ancestor i + 1 of the new slot, a ballot counts how far x rises.
// Slots are 1-based (children 2s, 2s + 1) on q1 = heap - 1, 16-byte aligned so that a child pair is one load.
// The heap array is identical to the literal loops' after every operation, so the tie-breaks of equal-frequency
// internal nodes - which decide the reference's code lengths - are preserved.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void warp_heap_offer(uint64_t* q1, int& size, uint64_t x, uint32_t lane) {
    const uint32_t s = (uint32_t)++size;                        // the new slot
    // lane i: ancestor i + 1 (s <= 256: at most 8); the clamping funnel shift gives 0 from lane 8 on and
    // q1[0] holds the smallest key, which never moves down
    const uint32_t anc = __funnelshift_rc(s, 0u, lane + 1);
    const uint64_t e = q1[anc];
    // cmp(x, ancestor) < 0: the ancestor moves down.  siftUp stops at the first ancestor that stays; ancestors
    // descend towards the root, so those that move are a prefix
    const uint32_t t = __popc(__ballot_sync(0xffffffffu, e > (x | 511ull)));
    if (lane <= t) q1[s >> lane] = lane < t ? e : x;
    __syncwarp();
}

__device__ __forceinline__ uint64_t warp_heap_poll(uint64_t* q1, int& size, uint32_t lane) {
    const uint32_t FULL = 0xffffffffu;
    const uint64_t result = q1[1];
    const uint32_t n = (uint32_t)--size;
    if (n == 0) return result;
    const uint64_t x = q1[n + 1];
    // Slots beyond the heap hold the largest key, so a missing child is never the smaller one and never smaller than
    // x: no bounds checks.  Slot n + 1 still holds x during this poll - harmless: as a right child it is preferred
    // only if the left one is larger than x (then x stays above both anyway), and x never sinks below itself.
    uint32_t p[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const ulonglong2 ch = *reinterpret_cast<const ulonglong2*>(q1 + 2 * (lane + 32 * i));   // children of slot lane + 32 i
        p[i] = __ballot_sync(FULL, key_gt(ch.x, ch.y));                                         // (slot 0: unused bit)
    }
    uint32_t s = 1;
#pragma unroll
    for (int L = 0; L < 7; ++L) {
        const uint32_t w = L < 5 ? p[0] : L == 5 ? p[1] : ((s & 32u) ? p[3] : p[2]);
        s = 2 * s + (__funnelshift_r(w, 0u, s) & 1u);           // shift amount taken mod 32
    }
    // s = 1 b0 b1 .. b6 in binary: the path slot of level L is its top L + 1 bits
    const uint32_t my_s = s >> (7 - min(lane, 7u));
    const uint32_t my_sn = lane < 7 ? s >> (6 - lane) : 257u;   // q1[257] is always the largest key
    const uint64_t c = q1[my_sn];
    // siftDown stops at the first level with cmp(x, child) <= 0; the entries along the path ascend, so the lanes that
    // sink are a prefix and their count is that level
    const uint32_t t = __popc(__ballot_sync(FULL, key_gt(x, c)));
    if (lane <= t) q1[my_s] = lane < t ? c : x;
    if (lane == 31) q1[n + 1] = ~0ull;                          // the vacated slot (all lanes have x: it fed the ballot)
    __syncwarp();
    return result;
}

// buildCodeLengths up to the tree, by all 32 lanes of a warp: leaves offered in ascending symbol order, then
// poll / poll / offer until one node is left.  Leaf s gets node id leaf_id[s] = its rank among the present symbols,
// internal nodes continue the numbering; parent[id] is the tree.  heapbuf: 258 words, 16-byte aligned.
__device__ __forceinline__ int warp_heap_replay(const uint32_t* hist, uint64_t* heapbuf, uint16_t* parent, uint16_t* leaf_id,
                                                int* root, uint32_t lane) {
    uint64_t* q1 = heapbuf;                                     // 1-based slots: q1[1..256]; q1[0], q1[257] never hold entries
    for (uint32_t i = lane; i < 258; i += 32) q1[i] = i ? ~0ull : 0ull;     // the largest key beyond the heap (warp_heap_poll), the smallest in q1[0] (warp_heap_offer)
    __syncwarp();
    int size = 0, n = 0;
    for (int s = 0; s < 256; ++s) {
        const uint32_t fr = hist[s];
        if (fr > 0) {
            if (lane == 0) leaf_id[s] = (uint16_t)n;
            warp_heap_offer(q1, size, ((uint64_t)fr << 18) | ((uint64_t)(s + 1) << 9) | (uint64_t)n, lane);
            ++n;
        }
    }
    const int nsym = n;
    while (size > 1) {
        const uint64_t l = warp_heap_poll(q1, size, lane);
        const uint64_t r = warp_heap_poll(q1, size, lane);
        if (lane < 2) parent[(lane ? r : l) & 511] = (uint16_t)n;
        warp_heap_offer(q1, size, (((l >> 18) + (r >> 18)) << 18) | (uint64_t)n, lane);
        ++n;
    }
    *root = n - 1;
    __syncwarp();
    return nsym;
}

// Block-wide exclusive scan helper for 64-bit values over `count` items stored in global memory
// (in place: in = per-item value, out = exclusive prefix).  Called by all CB_THREADS threads.
__device__ void block_scan_inplace_u64(uint64_t* data, uint32_t count, uint64_t* s_warp /*[9]*/) {
    const uint32_t t = threadIdx.x, lane = t & 31, wid = t >> 5;
    uint64_t carry = 0;
    for (uint32_t base = 0; base < count; base += CB_THREADS) {
        uint32_t i = base + t;
        uint64_t v = i < count ? data[i] : 0;
        uint64_t inc = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint64_t o = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += o;
        }
        if (lane == 31) s_warp[wid] = inc;
        __syncthreads();
        if (t == 0) {
            uint64_t a = 0;
            for (int w = 0; w < CB_THREADS / 32; ++w) { uint64_t x = s_warp[w]; s_warp[w] = a; a += x; }
            s_warp[8] = a;
        }
        __syncthreads();
        if (i < count) data[i] = carry + s_warp[wid] + inc - v;
        carry += s_warp[8];
        __syncthreads();
    }
}

// Shared memory of one codebook build (6.3 KiB).  The chained kernel overlays it on its histogram counters.
struct __align__(16) CbShared {
    uint64_t heap[258];
    uint64_t s_warp[9];
    uint32_t hist[256];
    int s_len[256];
    uint32_t lcount[34];
    uint32_t first[34];
    uint16_t parent[512];
    uint16_t leaf_id[256];
    int s_root, s_nsym, s_maxlen;
};

// The codebook of chunk k by the CB_THREADS threads of a CTA.  CHAIN: the segment histograms were written by other CTAs
// of the SAME grid (hist_chain_kernel): they are read through L2 (ld.global.cg), never through the non-coherent path.
// Returns (to every thread) the chunk's compressed size in bytes.
template <bool CHAIN>
__device__ __forceinline__ uint32_t codebook_body(CbShared& S, const uint32_t k, const uint32_t* __restrict__ seg_hist, uint32_t spc,
                                                  uint32_t* __restrict__ chunk_hist_out,
                                                  uint8_t* __restrict__ len_out, uint32_t* __restrict__ code_out,
                                                  uint64_t* __restrict__ chunk_bits, uint32_t* __restrict__ comp_size,
                                                  uint64_t* __restrict__ seg_bitoff, const uint8_t* __restrict__ fixed_len,
                                                  const uint32_t* __restrict__ direct_hist, int* status) {
    uint32_t* hist = S.hist; uint64_t* heap = S.heap; uint16_t* parent = S.parent; uint16_t* leaf_id = S.leaf_id;
    int* s_len = S.s_len; uint32_t* lcount = S.lcount; uint32_t* first = S.first; uint64_t* s_warp = S.s_warp;
    int& s_root = S.s_root; int& s_nsym = S.s_nsym; int& s_maxlen = S.s_maxlen;
    auto LD = [](const uint32_t* p) -> uint32_t { return CHAIN ? __ldcg(p) : __ldg(p); };
    uint32_t my_bytes = 0;

    const uint32_t t = threadIdx.x, lane = t & 31, wid = t >> 5;

    // 1. chunk histogram = sum of its segment histograms (or a caller-supplied histogram)
    uint32_t f = 0;
    if (direct_hist) {
        f = direct_hist[(size_t)k * 256 + t];
    } else {
        // a 16 MiB chunk has 300 segment histograms: 16 independent loads in flight per thread
        const uint32_t* sh = seg_hist + (size_t)k * spc * 256 + t;
        uint32_t s = 0;
        for (; s + 16 <= spc; s += 16) {
            uint32_t v[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = LD(sh + (size_t)(s + j) * 256);
#pragma unroll
            for (int j = 0; j < 16; ++j) f += v[j];
        }
        for (; s < spc; ++s) f += LD(sh + (size_t)s * 256);
    }
    hist[t] = f;
    if (chunk_hist_out) chunk_hist_out[(size_t)k * 256 + t] = f;
    if (t < 34) lcount[t] = 0;
    if (t == 0) s_maxlen = 0;
    __syncthreads();

    int mylen = 0;
    if (fixed_len) {
        mylen = fixed_len[t];
        if (f > 0 && mylen == 0) hz_set_status(status, HZ_ERR_BAD_LENGTHS);
    } else {
        // 2. the serial part: replay java.util.PriorityQueue (CanonicalHuffman.java:55-70)
#ifndef HZ_CB_PLAIN
        if (t < 32) {                           // warp 0, all lanes (warp_heap_replay)
            int root;
            const int nsym = warp_heap_replay(hist, heap, parent, leaf_id, &root, lane);
            if (t == 0) { s_nsym = nsym; s_root = root; }
        }
#else
        if (t == 0) {
            int size = 0, n = 0;
            for (int s = 0; s < 256; ++s) {
                uint32_t fr = hist[s];
                if (fr > 0) {
                    leaf_id[s] = (uint16_t)n;
                    heap_offer(heap, size, ((uint64_t)fr << 18) | ((uint64_t)(s + 1) << 9) | (uint64_t)n);
                    ++n;
                }
            }
            s_nsym = n;
            while (size > 1) {
                uint64_t l = heap_poll(heap, size);
                uint64_t r = heap_poll(heap, size);
                parent[l & 511] = (uint16_t)n;
                parent[r & 511] = (uint16_t)n;
                heap_offer(heap, size, (((l >> 18) + (r >> 18)) << 18) | (uint64_t)n);
                ++n;
            }
            s_root = n - 1;
        }
#endif
        __syncthreads();
        // 3. leaf depth = code length (extractLengths, :85-92); single symbol -> 1 (:35-45)
        if (f > 0) {
            if (s_nsym == 1) {
                mylen = 1;
            } else {
                int id = leaf_id[t], root = s_root, d = 0;
                while (id != root) { id = parent[id]; ++d; }
                mylen = d;
            }
        }
    }
    if (mylen > 32) {                           // the reference throws here (:107)
        hz_set_status(status, HZ_ERR_CODE_TOO_LONG);
        mylen = 0;
        s_maxlen = 99;
    }
    s_len[t] = mylen;
    if (mylen > 0) atomicAdd(&lcount[mylen], 1u);
    __syncthreads();
    const bool bad = s_maxlen == 99;
    if (bad) mylen = 0;

    // 4. canonical codes (generateCanonicalCodes, :99-132)
    if (t == 0) {
        uint32_t c = 0;
        first[0] = 0;
        for (int l = 1; l <= 32; ++l) {
            c = (c + (l > 1 ? lcount[l - 1] : 0u)) << 1;
            first[l] = c;
        }
    }
    __syncthreads();
    uint32_t mycode = 0;
    if (mylen > 0) {
        uint32_t rank = 0;
        for (uint32_t s = 0; s < t; ++s) rank += (s_len[s] == mylen);
        mycode = first[mylen] + rank;
    }
    len_out[(size_t)k * 256 + t] = (uint8_t)mylen;
    if (code_out) code_out[(size_t)k * 256 + t] = mycode;

    // 5. exact compressed size of the chunk: sum(freq * len) bits
    if (chunk_bits) {
        uint64_t b = (uint64_t)f * (uint32_t)mylen;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) b += __shfl_xor_sync(0xffffffffu, b, d);
        if (lane == 0) s_warp[wid] = b;
        __syncthreads();
        if (t == 0) {
            uint64_t a = 0;
            for (int w = 0; w < CB_THREADS / 32; ++w) a += s_warp[w];
            chunk_bits[k] = a;
            uint64_t bytes = (a + 7) >> 3;
            if (bytes > 0x7fffffffull) { hz_set_status(status, HZ_ERR_OUT_TOO_SMALL); bytes = 0; }
            comp_size[k] = (uint32_t)bytes;
            s_warp[8] = bytes;
        }
        __syncthreads();
        my_bytes = (uint32_t)s_warp[8];
        __syncthreads();
    }

    // 6. bit offset of every segment inside the chunk's stream
    if (seg_bitoff && seg_hist) {
        uint64_t* so = seg_bitoff + (size_t)k * spc;
        uint32_t l8[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) l8[j] = bad ? 0u : (uint32_t)s_len[lane + 32 * j];
        // four segments per warp and iteration: 32 independent loads in flight per lane, the four reductions interleaved
        for (uint32_t s0 = wid * 4; s0 < spc; s0 += CB_THREADS / 32 * 4) {
            uint32_t v[4][8];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const uint32_t* sh = seg_hist + ((size_t)k * spc + min(s0 + q, spc - 1)) * 256 + lane;
#pragma unroll
                for (int j = 0; j < 8; ++j) v[q][j] = LD(sh + 32 * j);
            }
            uint64_t b[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                b[q] = 0;
#pragma unroll
                for (int j = 0; j < 8; ++j) b[q] += (uint64_t)v[q][j] * l8[j];
            }
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) {
#pragma unroll
                for (int q = 0; q < 4; ++q) b[q] += __shfl_xor_sync(0xffffffffu, b[q], d);
            }
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (lane == (uint32_t)q && s0 + q < spc) so[s0 + q] = b[q];
        }
        __syncthreads();
        block_scan_inplace_u64(so, spc, s_warp);
    }
    return my_bytes;
}

__global__ void __launch_bounds__(CB_THREADS)
codebook_kernel(const uint32_t* __restrict__ seg_hist, uint32_t spc, uint32_t* __restrict__ chunk_hist_out,
                uint8_t* __restrict__ len_out, uint32_t* __restrict__ code_out,
                uint64_t* __restrict__ chunk_bits, uint32_t* __restrict__ comp_size,
                uint64_t* __restrict__ seg_bitoff, const uint8_t* __restrict__ fixed_len,
                const uint32_t* __restrict__ direct_hist, int* status) {
    __shared__ CbShared S;
    codebook_body<false>(S, blockIdx.x, seg_hist, spc, chunk_hist_out, len_out, code_out, chunk_bits, comp_size, seg_bitoff,
                         fixed_len, direct_hist, status);
}

// ---------------------------------------------------------------------------------------------
// Chained histogram -> codebook -> offsets (large streams of large chunks; hz_api.cu: encode_device).
// The codebook stage is 0.16 ms of pure latency (one warp replays one chunk's heap) whatever the chunk count; as a
// kernel of its own it sits between the histogram and the encoder.  Here the histogram CTAs take their ranges from a
// ticket (so that start order == range order), and the CTA that completes a chunk's histogram (a counter per chunk)
// builds that chunk's codebook on the spot, obtains the chunk's output offset by a decoupled look-back over the
// chunks before it (prefix[k-1]: its predecessor's tail has a smaller ticket, so it is running or done) and raises
// ready[k].  Every CTA issues griddepcontrol.launch_dependents at its start: the encoder, launched with programmatic
// stream serialization, becomes resident as histogram CTAs retire and its CTAs wait for ready[chunk] - the last
// chunks' codebooks are built while the first chunks are being encoded.  All waits are bounded (a fault latches
// HZ_ERR_CUDA instead of hanging the device).
// ---------------------------------------------------------------------------------------------
#define HZ_CHAIN_SPIN_MAX (1u << 22)
__device__ __forceinline__ uint64_t chain_ld_acquire(const uint64_t* p) {
    uint64_t v; asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ void chain_st_release(uint64_t* p, uint64_t v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__global__ void __launch_bounds__(HZ_THREADS, 6)
hist_chain_kernel(const uint8_t* __restrict__ in, uint64_t n, uint32_t chunk_bytes, uint32_t spc, uint32_t mult, uint32_t K,
                  uint32_t* __restrict__ seg_hist, HzChain c, uint32_t* __restrict__ chunk_hist_out,
                  uint8_t* __restrict__ len_out, uint32_t* __restrict__ code_out, uint64_t* __restrict__ chunk_bits,
                  uint32_t* __restrict__ comp_size, uint64_t* __restrict__ comp_off, uint64_t* __restrict__ seg_bitoff,
                  int* status) {
    __shared__ __align__(16) uint32_t h[256 * 32];
    __shared__ uint32_t s_bid, s_last;
    static_assert(sizeof(CbShared) <= sizeof(uint32_t) * 256 * 32, "the codebook's shared memory overlays the counters");
    static_assert(HZ_THREADS == CB_THREADS, "the tail runs with the histogram's CTA");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const uint32_t t = threadIdx.x;
    if (t == 0) s_bid = atomicAdd(c.ticket, 1u);
    __syncthreads();
    const uint32_t bid = s_bid;
    const uint32_t rpc = (spc + mult - 1) / mult;
    const uint32_t k = bid / rpc;
    hist_range_lanes(h, in, n, chunk_bytes, spc, mult, seg_hist, bid);
    __threadfence();                                      // this CTA's bins are visible device-wide ...
    __syncthreads();
    if (t == 0) s_last = atomicAdd(c.done + k, 1u) == rpc - 1 ? 1u : 0u;   // ... before it counts as done
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // ---- tail: this CTA completed chunk k's histogram ----------------------------------------------------
    CbShared& S = *reinterpret_cast<CbShared*>(h);
    const uint32_t bytes = codebook_body<true>(S, k, seg_hist, spc, chunk_hist_out, len_out, code_out, chunk_bits, comp_size,
                                               seg_bitoff, nullptr, nullptr, status);
    __threadfence();                                      // lengths, codes, segment offsets: visible before ready[k]
    __syncthreads();
    // chunk offsets: DECOUPLED look-back by warp 0.  prefix[k] carries a state in its top two bits: AGGREGATE (this chunk's
    // size alone, published at once) or PREFIX (payload bytes of chunks 0..k).  A tail sums the aggregates behind it, 32
    // records at a time, until it meets a PREFIX - it never waits for its predecessor's look-back, only for its size.
    // (A chain of prefix[k-1] -> prefix[k] hops costs ~2 us per chunk in SERIES: fine while tails finish 3 us apart, as
    // with 16 MiB chunks, a bottleneck at 8 MiB and below - 4 MiB chunks: 3.30 -> 3.94 ms per 4 GiB.)
    if (t < 32) {
        const uint64_t ST_A = 1ull << 62, ST_P = 2ull << 62, VAL = (1ull << 62) - 1;
        if (t == 0) chain_st_release(c.prefix + k, ST_A | bytes);
        uint64_t off = 0;
        int base = (int)k - 1;
        uint32_t spins = 0, ns = 32;
        while (base >= 0) {
            const int j = base - (int)t;
            const uint64_t v = j >= 0 ? chain_ld_acquire(c.prefix + j) : ST_P;        // before chunk 0: a prefix of 0
            const uint32_t m_empty = __ballot_sync(0xffffffffu, (v >> 62) == 0);
            const uint32_t m_pref = __ballot_sync(0xffffffffu, (v >> 62) >= 2);
            const uint32_t first_p = m_pref ? (uint32_t)__ffs(m_pref) - 1 : 32u;       // nearest record that is a PREFIX
            const uint32_t need = first_p >= 31 ? 0xffffffffu : ((2u << first_p) - 1); // records up to and including it
            if (m_empty & need) {                         // one of them is not published yet
                if (++spins > HZ_CHAIN_SPIN_MAX) { hz_set_status(status, HZ_ERR_CUDA); break; }
                __nanosleep(ns); if (ns < 512) ns += ns;
                continue;
            }
            uint64_t x = ((need >> t) & 1) ? (v & VAL) : 0;
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) x += __shfl_xor_sync(0xffffffffu, x, d);
            off += x;
            if (m_pref) break;
            base -= 32;                                   // 32 aggregates, no prefix among them: further back
        }
        if (t == 0) {
            comp_off[k] = off;
            if (k == K - 1) comp_off[K] = off + bytes;
            chain_st_release(c.prefix + k, ST_P | (off + bytes));
            __threadfence();
            chain_st_release(c.ready + k, 1ull);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Many small chunks: ONE WARP per chunk (8 chunks per CTA), each replaying its chunk's heap with
// warp_heap_replay: an SM keeps ~40 heaps in flight instead of the 8 of one CTA per chunk, which is what
// bounds the codebook stage when a stream is cut into thousands of chunks (256 KiB chunks: 4 k per GiB).
// Same results as codebook_kernel; used when spc <= 32 (the segment offsets are scanned by one warp).
// ---------------------------------------------------------------------------------------------
#define CBW_WARPS 8
struct __align__(16) CbWarp {
    uint64_t heap[258];                 // 1-based slots for warp_heap_replay; sizeof(CbWarp) is a multiple of 16
    uint32_t hist[256];
    uint16_t parent[512];
    uint16_t leaf_id[256];
    uint8_t len[256];
    uint32_t lcount[34];
    uint32_t first[34];
};

__global__ void __launch_bounds__(CBW_WARPS * 32)
codebook_warp_kernel(const uint32_t* __restrict__ seg_hist, uint32_t spc, uint32_t K, uint32_t* __restrict__ chunk_hist_out,
                     uint8_t* len_out, uint32_t* __restrict__ code_out,
                     uint64_t* __restrict__ chunk_bits, uint32_t* __restrict__ comp_size,
                     uint64_t* __restrict__ seg_bitoff, const uint8_t* __restrict__ fixed_len,
                     const uint32_t* __restrict__ direct_hist, int lens_ready, int* status) {
    __shared__ CbWarp S[CBW_WARPS];
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t k = blockIdx.x * CBW_WARPS + wid;
    if (k >= K) return;
    CbWarp& W = S[wid];
    const uint32_t FULL = 0xffffffffu;

    // 1. chunk histogram (lane handles symbols lane + 32 j)
    uint32_t f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const uint32_t sym = lane + 32 * j;
        uint32_t a = 0;
        if (direct_hist) a = direct_hist[(size_t)k * 256 + sym];
        else for (uint32_t s = 0; s < spc; ++s) a += seg_hist[((size_t)k * spc + s) * 256 + sym];
        f[j] = a;
        W.hist[sym] = a;
        if (chunk_hist_out) chunk_hist_out[(size_t)k * 256 + sym] = a;
    }
    if (lane < 34) W.lcount[lane] = 0;
    if (lane == 0) W.lcount[32] = W.lcount[33] = 0;
    __syncwarp();

    uint32_t mylen[8];
    bool bad = false;
    if (fixed_len) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            mylen[j] = fixed_len[lane + 32 * j];
            if (f[j] > 0 && mylen[j] == 0) hz_set_status(status, HZ_ERR_BAD_LENGTHS);
        }
    } else if (lens_ready == 1) {
        // the heap replay was done by codebook_lane_kernel (32 chunks per warp): leaf depths are in len_out
#pragma unroll
        for (int j = 0; j < 8; ++j) mylen[j] = len_out[(size_t)k * 256 + lane + 32 * j];
    } else {
        // 2. the serial part: replay java.util.PriorityQueue (CanonicalHuffman.java:55-70)
        int nsym = 0, root = 0;
#ifndef HZ_CB_PLAIN
        if (lens_ready == 0) {
            nsym = warp_heap_replay(W.hist, W.heap, W.parent, W.leaf_id, &root, lane);
        } else                                  // lens_ready == 2: the one-lane loops (developer knob HZ_CODEBOOK_REPLAY=lane0)
#endif
        if (lane == 0) {
            int size = 0, n = 0;
            for (int s = 0; s < 256; ++s) {
                const uint32_t fr = W.hist[s];
                if (fr > 0) {
                    W.leaf_id[s] = (uint16_t)n;
                    heap_offer(W.heap, size, ((uint64_t)fr << 18) | ((uint64_t)(s + 1) << 9) | (uint64_t)n);
                    ++n;
                }
            }
            nsym = n;
            while (size > 1) {
                const uint64_t l = heap_poll(W.heap, size);
                const uint64_t r = heap_poll(W.heap, size);
                W.parent[l & 511] = (uint16_t)n;
                W.parent[r & 511] = (uint16_t)n;
                heap_offer(W.heap, size, (((l >> 18) + (r >> 18)) << 18) | (uint64_t)n);
                ++n;
            }
            root = n - 1;
        }
        nsym = __shfl_sync(FULL, nsym, 0);
        root = __shfl_sync(FULL, root, 0);
        __syncwarp();
        // 3. leaf depth = code length (extractLengths, :85-92); single symbol -> 1 (:35-45)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            uint32_t d = 0;
            if (f[j] > 0) {
                if (nsym == 1) d = 1;
                else { int id = W.leaf_id[lane + 32 * j]; while (id != root) { id = W.parent[id]; ++d; } }
            }
            mylen[j] = d;
        }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) bad |= mylen[j] > 32;
    bad = __any_sync(FULL, bad);
    if (bad) {                                  // the reference throws here (:107)
        if (lane == 0) hz_set_status(status, HZ_ERR_CODE_TOO_LONG);
#pragma unroll
        for (int j = 0; j < 8; ++j) mylen[j] = 0;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        W.len[lane + 32 * j] = (uint8_t)mylen[j];
        if (mylen[j] > 0) atomicAdd(&W.lcount[mylen[j]], 1u);
    }
    __syncwarp();
    // 4. canonical codes (generateCanonicalCodes, :99-132): first code of every length, then the
    //    symbols of a length in increasing symbol order (lane 0 walks the alphabet once)
    if (lane == 0) {
        uint32_t c = 0;
        W.first[0] = 0;
        for (int l = 1; l <= 32; ++l) { c = (c + (l > 1 ? W.lcount[l - 1] : 0u)) << 1; W.first[l] = c; }
        for (int s = 0; s < 256; ++s) {
            const uint32_t l = W.len[s];
            W.hist[s] = l ? W.first[l]++ : 0u;   // hist is no longer needed (f[] holds the counts): reuse for the codes
        }
    }
    __syncwarp();
    uint64_t bits = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const uint32_t sym = lane + 32 * j;
        len_out[(size_t)k * 256 + sym] = (uint8_t)mylen[j];
        if (code_out) code_out[(size_t)k * 256 + sym] = W.hist[sym];
        bits += (uint64_t)f[j] * mylen[j];
    }
    // 5. exact compressed size of the chunk: sum(freq * len) bits
    if (chunk_bits) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) bits += __shfl_xor_sync(FULL, bits, d);
        if (lane == 0) {
            chunk_bits[k] = bits;
            uint64_t bytes = (bits + 7) >> 3;
            if (bytes > 0x7fffffffull) { hz_set_status(status, HZ_ERR_OUT_TOO_SMALL); bytes = 0; }
            comp_size[k] = (uint32_t)bytes;
        }
    }
    // 6. bit offset of every segment inside the chunk's stream (spc <= 32: one lane per segment)
    if (seg_bitoff && seg_hist) {
        uint64_t mine = 0;
        for (uint32_t s = 0; s < spc; ++s) {
            const uint32_t* sh = seg_hist + ((size_t)k * spc + s) * 256;
            uint64_t b = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) b += (uint64_t)sh[lane + 32 * j] * mylen[j];
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) b += __shfl_xor_sync(FULL, b, d);
            if (lane == s) mine = b;
        }
        uint64_t inc = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint64_t o = __shfl_up_sync(FULL, inc, d);
            if (lane >= (uint32_t)d) inc += o;
        }
        if (lane < spc) seg_bitoff[(size_t)k * spc + lane] = inc - mine;
    }
}

// ---------------------------------------------------------------------------------------------
// Thousands of chunks: ONE LANE per chunk for the serial heap replay.  With one active lane per warp the
// replay (about 10^5 issue slots per chunk) costs as many issue slots as 32 chunks in lock step do, and it
// bounded the encode stage of 64 KiB chunks (1.7 ms of 2.7 ms per GiB).  Here every lane of a warp replays
// the JDK heap of its own chunk; the lanes diverge only in the sift depths.  The kernel is latency bound
// (a dependent chain of shared-memory loads and compares per heap level), so what counts is how many warps
// an SM holds, i.e. shared memory per chunk - 1.75 KiB:
//   keys  u32 [256][32]  (freq << 9) | (symbol + 1), internal nodes carry field 0: comparing keys IS
//                        HuffmanNode.compareTo (two internal nodes of equal frequency compare equal).  Needs
//                        freq < 2^23, guaranteed by the caller (chunks of at most 32 segments = 1.75 MiB);
//   ids   u8  [256][32]  index m of the INTERNAL node in a heap slot (a leaf's identity is in its key);
//   par   u8  [512][32]  parent (always an internal node, m = 0..254; 0xFF = symbol absent) of leaf s at
//                        [s], of internal node m at [256 + m];
// all [slot][lane], so a lane stays in its own bank whatever the slots.  The histogram columns are staged
// through the par+ids area in two halves (coalesced loads); after the merge loop the keys area holds the
// internal nodes' depths, filled top-down (a parent's index is larger than its children's).
// Output: the code LENGTHS (leaf depths; a single symbol gets 1) in len_out;
// codebook_warp_kernel(lens_ready) derives codes, sizes and segment offsets from them.
// ---------------------------------------------------------------------------------------------
#define CBL_WARPS 4
#define CBL_KEYS_BYTES (256 * 32 * 4)
#define CBL_PAR_BYTES (512 * 32)
#define CBL_IDS_BYTES (256 * 32)
#define CBL_WARP_BYTES (CBL_KEYS_BYTES + CBL_PAR_BYTES + CBL_IDS_BYTES)     // 56 KiB
#define CBL_SMEM (CBL_WARPS * CBL_WARP_BYTES)
#define CBL_HS 33                                                           // staging stride (words) of one symbol's 32 columns
static_assert(128 * CBL_HS * 4 <= CBL_PAR_BYTES + CBL_IDS_BYTES, "half a histogram is staged in the par+ids area");
static_assert(CBL_SMEM <= 227 * 1024, "one CTA of CBL_WARPS warps per SM");

__global__ void __launch_bounds__(CBL_WARPS * 32)
codebook_lane_kernel(const uint32_t* __restrict__ seg_hist, uint32_t spc, uint32_t K, uint8_t* __restrict__ len_out) {
    extern __shared__ __align__(16) uint8_t cbl_smem[];
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint8_t* wsm = cbl_smem + wid * CBL_WARP_BYTES;
    uint32_t* keys = reinterpret_cast<uint32_t*>(wsm) + lane;                           // slot i at keys[i * 32]
    uint8_t* par = wsm + CBL_KEYS_BYTES + lane;                                         // node i at par[i * 32]
    uint8_t* ids = wsm + CBL_KEYS_BYTES + CBL_PAR_BYTES + lane;                         // slot i at ids[i * 32]
    uint32_t* hs = reinterpret_cast<uint32_t*>(wsm + CBL_KEYS_BYTES);                   // staging [sym * CBL_HS + chunk]
    uint8_t* dep = wsm + lane;                                                          // internal node m at dep[m * 32]
    const uint32_t k0 = (blockIdx.x * CBL_WARPS + wid) * 32;
    if (k0 >= K) return;

    // 1 + 2. leaves in ascending symbol order (CanonicalHuffman.java:56-62), offer = siftUp; the chunk histograms of
    // the warp's 32 chunks come in two halves of 128 symbols (coalesced: lane = symbol lane + 32 j)
    int size = 0, n = 0, only = 0;
    for (int half = 0; half < 2; ++half) {
        for (uint32_t c = 0; c < 32; ++c) {
            const uint32_t k = k0 + c;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t sl = lane + 32 * j;                // symbol within the half
                uint32_t a = 0;
                if (k < K) for (uint32_t s = 0; s < spc; ++s) a += seg_hist[((size_t)k * spc + s) * 256 + half * 128 + sl];
                hs[sl * CBL_HS + c] = a;
            }
        }
        __syncwarp();
        for (int sl = 0; sl < 128; ++sl) {
            const uint32_t fr = hs[sl * CBL_HS + lane];
            if (fr > 0) {
                const int s = half * 128 + sl;
                const uint32_t x = (fr << 9) | (uint32_t)(s + 1);
                int i = size++;
                while (i > 0) {
                    const int p = (i - 1) >> 1;
                    const uint32_t e = keys[p * 32];
                    if (!(e > x)) break;                          // cmp(x, parent) >= 0
                    keys[i * 32] = e;
                    i = p;
                }
                keys[i * 32] = x;
                ++n; only = s;
            }
        }
        __syncwarp();                                             // every lane has consumed its staged column
    }
    for (int s = 0; s < 256; ++s) par[s * 32] = 0xFFu;           // absent symbols keep this
    // 3. merge loop (CanonicalHuffman.java:64-70): two polls (move last to root + siftDown), one offer (siftUp)
    int m = 0;                                                    // next internal node
    while (size > 1) {
        uint32_t fsum = 0;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const uint32_t top = keys[0];
            const uint32_t topid = ids[0];
            const int cnt = --size;
            if (cnt > 0) {
                const uint32_t x = keys[cnt * 32];
                const uint32_t xid = ids[cnt * 32];
                int i = 0;
                const int halfn = cnt >> 1;
                while (i < halfn) {
                    int child = 2 * i + 1;
                    uint32_t c = keys[child * 32];
                    const int right = child + 1;
                    if (right < cnt) {
                        const uint32_t r = keys[right * 32];
                        if (c > r) { c = r; child = right; }      // cmp(left, right) > 0
                    }
                    if (!(x > c)) break;                          // cmp(x, child) <= 0
                    keys[i * 32] = c;
                    ids[i * 32] = ids[child * 32];
                    i = child;
                }
                keys[i * 32] = x;
                ids[i * 32] = (uint8_t)xid;
            }
            const uint32_t sf = top & 511u;                       // 0: internal node topid, else leaf symbol sf - 1
            par[(sf ? sf - 1 : 256 + topid) * 32] = (uint8_t)m;
            fsum += top >> 9;
        }
        const uint32_t x = fsum << 9;
        int i = size++;
        while (i > 0) {
            const int p = (i - 1) >> 1;
            const uint32_t e = keys[p * 32];
            if (!(e > x)) break;
            keys[i * 32] = e;
            ids[i * 32] = ids[p * 32];
            i = p;
        }
        keys[i * 32] = x;
        ids[i * 32] = (uint8_t)m;
        ++m;
    }
    __syncwarp();                                   // the heaps are dead: their memory now holds node depths
    // 4. depths top-down (extractLengths, :85-92); internal node m - 1 is the root
    if (n >= 2) {
        dep[(m - 1) * 32] = 0;
        for (int j = m - 2; j >= 0; --j) dep[j * 32] = (uint8_t)(dep[(uint32_t)par[(256 + j) * 32] * 32] + 1u);
    }
    const uint32_t k = k0 + lane;
    if (k < K) {
        uint32_t* lo = reinterpret_cast<uint32_t*>(len_out + (size_t)k * 256);
        for (int s4 = 0; s4 < 64; ++s4) {
            uint32_t w = 0;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const int s = s4 * 4 + b;
                uint32_t l = 0;
                if (n == 1) l = (s == only) ? 1u : 0u;
                else if (n >= 2) {
                    const uint32_t p = par[s * 32];
                    if (p != 0xFFu) { l = dep[p * 32] + 1u; if (l > 255u) l = 255u; }     // > 32 is an error downstream
                }
                w |= l << (8 * b);
            }
            lo[s4] = w;
        }
    }
}

// comp_off[k] = exclusive prefix sum of comp_size (K+1 entries), one CTA of 1024 threads:
// each thread sums a contiguous slice, the slice totals are scanned, then prefixes are written.
__global__ void __launch_bounds__(1024)
chunk_offsets_kernel(const uint32_t* __restrict__ comp_size, uint32_t K, uint64_t* __restrict__ comp_off) {
    __shared__ uint64_t part[1024];
    const uint32_t t = threadIdx.x;
    const uint32_t per = (K + 1023) / 1024;
    const uint32_t lo = t * per, hi = min(K, lo + per);
    uint64_t s = 0;
    for (uint32_t i = lo; i < hi; ++i) s += comp_size[i];
    part[t] = s;
    __syncthreads();
    if (t < 32) {   // warp 0 scans 1024 partials, 32 per lane
        uint64_t loc[32]; uint64_t a = 0;
#pragma unroll
        for (int j = 0; j < 32; ++j) { loc[j] = a; a += part[t * 32 + j]; }
        uint64_t inc = a;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint64_t o = __shfl_up_sync(0xffffffffu, inc, d);
            if (t >= (uint32_t)d) inc += o;
        }
        uint64_t base = inc - a;
#pragma unroll
        for (int j = 0; j < 32; ++j) part[t * 32 + j] = base + loc[j];
        if (t == 31) comp_off[K] = inc;
    }
    __syncthreads();
    uint64_t a = part[t];
    for (uint32_t i = lo; i < hi; ++i) { comp_off[i] = a; a += comp_size[i]; }
}

// CanonicalHuffman.generateCanonicalCodesFromLengths (core/CanonicalHuffman.java:141-146)
__global__ void __launch_bounds__(CB_THREADS)
codes_from_lengths_kernel(const uint8_t* __restrict__ len_in, uint32_t* __restrict__ code_out, int* status) {
    __shared__ int s_len[256];
    __shared__ uint32_t lcount[34];
    __shared__ uint32_t first[34];
    const uint32_t k = blockIdx.x, t = threadIdx.x;
    int mylen = len_in[(size_t)k * 256 + t];
    if (t < 34) lcount[t] = 0;
    __syncthreads();
    if (mylen > 32) { hz_set_status(status, HZ_ERR_BAD_LENGTHS); mylen = 0; }
    s_len[t] = mylen;
    if (mylen > 0) atomicAdd(&lcount[mylen], 1u);
    __syncthreads();
    if (t == 0) {
        uint32_t c = 0;
        first[0] = 0;
        for (int l = 1; l <= 32; ++l) { c = (c + (l > 1 ? lcount[l - 1] : 0u)) << 1; first[l] = c; }
    }
    __syncthreads();
    uint32_t mycode = 0;
    if (mylen > 0) {
        uint32_t rank = 0;
        for (uint32_t s = 0; s < t; ++s) rank += (s_len[s] == mylen);
        mycode = first[mylen] + rank;
    }
    code_out[(size_t)k * 256 + t] = mycode;
}

int hzk_codebook(hz_ctx* ctx, const uint32_t* d_seg_hist, uint32_t spc, uint32_t K,
                 uint32_t* d_chunk_hist, uint8_t* d_len, uint32_t* d_code, uint64_t* d_chunk_bits,
                 uint32_t* d_comp_size, uint64_t* d_comp_off, uint64_t* d_seg_bitoff,
                 const uint8_t* d_fixed_len256) {
    if (K == 0) {
        if (d_comp_off) HZ_CUDA(ctx, cudaMemsetAsync(d_comp_off, 0, sizeof(uint64_t), ctx->stream));
        return HZ_OK;
    }
    // spc == 0 means d_seg_hist is a caller-supplied K x 256 chunk histogram
    const uint32_t* direct = spc == 0 ? d_seg_hist : nullptr;
    // thousands of small chunks: warp per chunk (more serial heaps in flight per SM)
    if (K >= 1024 && spc <= 32) {
        int lens_ready = 0;
        // lane-per-chunk replay (segment histograms only: its 32-bit heap keys need freq < 2^23, and spc <= 32
        // bounds a chunk at 1.75 MiB): a warp of 32 chunks takes ~1.5x as long as one chunk on one lane, so it
        // pays once the warp-per-chunk kernel needs more than one wave of ~40 warps per SM
        bool lanes = !d_fixed_len256 && spc >= 1 && K >= 8192;
        if (ctx->knobs.codebook) lanes = !d_fixed_len256 && spc >= 1 && ctx->knobs.codebook == 2;   // developer knob: warp | lane
        if (lanes) {
            if (!ctx->attr_codebook) {
                HZ_CUDA(ctx, cudaFuncSetAttribute(codebook_lane_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CBL_SMEM));
                ctx->attr_codebook = true;
            }
            HZ_LAUNCH(ctx, "codebook_heap", codebook_lane_kernel, (K + CBL_WARPS * 32 - 1) / (CBL_WARPS * 32), CBL_WARPS * 32, CBL_SMEM,
                      d_seg_hist, spc, K, d_len);
            lens_ready = 1;
        }
        // replay inside the warp kernel: by all 32 lanes (warp_heap_replay) - less latency per chunk while the warps do
        // not fill the GPU (1 MiB chunks, K = 1,024 per GiB: 0.29 -> 0.195 ms) and fewer issue slots when they do
        // (256 KiB chunks, K = 4,096: 0.557 -> 0.470 ms); mode 2 keeps the one-lane loops for A/B runs
        if (!lens_ready) {
            bool coop = true;
            if (ctx->knobs.codebook_lane0) coop = false;     // developer knob: warp | lane0
            if (!coop) lens_ready = 2;
        }
        HZ_LAUNCH(ctx, "codebook", codebook_warp_kernel, (K + CBW_WARPS - 1) / CBW_WARPS, CBW_WARPS * 32, 0,
                  spc == 0 ? nullptr : d_seg_hist, spc, K, d_chunk_hist, d_len, d_code, d_chunk_bits,
                  d_comp_size, d_seg_bitoff, d_fixed_len256, direct, lens_ready, ctx->d_status);
    } else {
        HZ_LAUNCH(ctx, "codebook", codebook_kernel, K, CB_THREADS, 0,
                  spc == 0 ? nullptr : d_seg_hist, spc, d_chunk_hist, d_len, d_code, d_chunk_bits,
                  d_comp_size, d_seg_bitoff, d_fixed_len256, direct, ctx->d_status);
    }
    if (d_comp_off)
        HZ_LAUNCH(ctx, "chunk_offsets", chunk_offsets_kernel, 1, 1024, 0, d_comp_size, K, d_comp_off);
    return HZ_OK;
}

// histogram + codebooks + chunk offsets of K chunks in ONE launch (hist_chain_kernel); the encoder that follows waits
// for c.ready[chunk]
int hzk_hist_codebook_chain(hz_ctx* ctx, const uint8_t* d_in, uint64_t n, uint32_t chunk_bytes, uint32_t K, uint32_t* d_seg_hist,
                            const HzChain& c, uint32_t* d_chunk_hist, uint8_t* d_len, uint32_t* d_code, uint64_t* d_chunk_bits,
                            uint32_t* d_comp_size, uint64_t* d_comp_off, uint64_t* d_seg_bitoff) {
    const uint32_t spc = (chunk_bytes + HZ_SEG_BYTES - 1) / HZ_SEG_BYTES;
    const uint32_t mult = hz_range_mult(spc, ctx->knobs.range_mult);
    const uint32_t rpc = (spc + mult - 1) / mult;
    const uint64_t grid = (uint64_t)K * rpc;
    if (grid > 0x7fffffffull) return hz_fail(ctx, HZ_ERR_ARG, "too many ranges");
    HZ_LAUNCH(ctx, "hist_codebook_chain", hist_chain_kernel, (unsigned)grid, HZ_THREADS, 0, d_in, n, chunk_bytes, spc, mult, K,
              d_seg_hist, c, d_chunk_hist, d_len, d_code, d_chunk_bits, d_comp_size, d_comp_off, d_seg_bitoff, ctx->d_status);
    return HZ_OK;
}

int hzk_codes_from_lengths(hz_ctx* ctx, const uint8_t* d_len, uint32_t K, uint32_t* d_code) {
    if (K == 0) return HZ_OK;
    HZ_LAUNCH(ctx, "codes_from_lengths", codes_from_lengths_kernel, K, CB_THREADS, 0, d_len, d_code, ctx->d_status);
    return HZ_OK;
}
