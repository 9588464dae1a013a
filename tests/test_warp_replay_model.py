"""A lane-by-lane Python model of warp_heap_replay (csrc/hz_codebook.cu) against the oracle's java.util.PriorityQueue.

The codebook kernels do not run PriorityQueue.siftDown level by level: every lane derives "right child is the
smaller one" bits for four inner slots, the min-child path follows from the bits, lane L compares the path entry of
level L + 1 with the sifted element and a popcount of the ballot is the level it stops at; slots beyond the heap
hold the largest key instead of bounds checks and the vacated slot still holds the sifted element during a poll.
The CUDA code itself is checked on the GPU (tests/test_gpu_parity.py::test_warp_replay_heavy_ties_many_chunks);
this model pins the ALGORITHM on the CPU: same code lengths as the oracle where the reference's tie-breaks
(CanonicalHuffman.java:55-80, HuffmanNode.java:52-58) decide them, and the same heap array as the literal sift
loops after every operation.
"""
import numpy as np
import pytest

import orc

MAXKEY = (1 << 64) - 1


def key_gt(a, b):                       # HuffmanNode.compareTo(a, b) > 0 on (freq << 18 | (symbol + 1) << 9 | id)
    return a > (b | 511)


def warp_offer(q1, size, x):
    s = size + 1
    anc = [(s >> (lane + 1)) if lane < 8 else 0 for lane in range(32)]
    e = [q1[a] for a in anc]                                        # q1[0] holds the smallest key
    ballot = [ev > (x | 511) for ev in e]
    t = sum(ballot)
    assert ballot == [True] * t + [False] * (32 - t), "the ancestors that move down must be a prefix"
    for lane in range(32):
        if lane <= t:
            q1[s >> lane] = e[lane] if lane < t else x
    return s


def warp_poll(q1, size):
    result = q1[1]
    n = size - 1
    if n == 0:
        return result, n
    x = q1[n + 1]
    p = [0, 0, 0, 0]
    for i in range(4):
        for lane in range(32):
            sl = lane + 32 * i
            if key_gt(q1[2 * sl], q1[2 * sl + 1]):
                p[i] |= 1 << lane
    s = 1
    for L in range(7):
        w = p[0] if L < 5 else p[1] if L == 5 else (p[3] if s & 32 else p[2])
        s = 2 * s + ((w >> (s & 31)) & 1)
    my_s = [s >> (7 - min(lane, 7)) for lane in range(32)]
    my_sn = [(s >> (6 - lane)) if lane < 7 else 257 for lane in range(32)]
    c = [q1[i] for i in my_sn]
    ballot = [key_gt(x, cv) for cv in c]
    t = sum(ballot)
    assert ballot == [True] * t + [False] * (32 - t), "the lanes that sink must be a prefix"
    for lane in range(32):
        if lane <= t:
            q1[my_s[lane]] = c[lane] if lane < t else x
    q1[n + 1] = MAXKEY
    return result, n


def plain_offer(q, size, x):            # PriorityQueue.siftUp
    k = size
    while k > 0:
        parent = (k - 1) >> 1
        if not key_gt(q[parent], x):
            break
        q[k] = q[parent]
        k = parent
    q[k] = x
    return size + 1


def plain_poll(q, size):                # PriorityQueue.poll + siftDown
    result = q[0]
    n = size - 1
    if n > 0:
        x = q[n]
        k, half = 0, n >> 1
        while k < half:
            child = 2 * k + 1
            if child + 1 < n and key_gt(q[child], q[child + 1]):
                child += 1
            if not key_gt(x, q[child]):
                break
            q[k] = q[child]
            k = child
        q[k] = x
    return result, n


def replay_lengths(hist, check_state=True):
    q1 = [MAXKEY] * 258
    q1[0] = 0
    q = [0] * 256
    size = psize = 0
    n = 0
    parent, leaf_id = {}, {}

    def same():
        assert size == psize and q1[1:size + 1] == q[:size], "heap differs from the literal loops"
        # (a poll that empties the heap returns early and leaves slot 1 as it was; the next offer overwrites it)
        assert all(v == MAXKEY for v in q1[max(size, 1) + 1:]), "slots beyond the heap must hold the largest key"

    for s in range(256):
        if hist[s]:
            leaf_id[s] = n
            e = (int(hist[s]) << 18) | ((s + 1) << 9) | n
            size = warp_offer(q1, size, e)
            if check_state:
                psize = plain_offer(q, psize, e); same()
            n += 1
    nsym = n
    while size > 1:
        l, size = warp_poll(q1, size)
        if check_state:
            pl, psize = plain_poll(q, psize); assert pl == l; same()
        r, size = warp_poll(q1, size)
        if check_state:
            pr, psize = plain_poll(q, psize); assert pr == r; same()
        parent[l & 511] = parent[r & 511] = n
        e = (((l >> 18) + (r >> 18)) << 18) | n
        size = warp_offer(q1, size, e)
        if check_state:
            psize = plain_offer(q, psize, e); same()
        n += 1
    root = n - 1
    lens = np.zeros(256, dtype=np.int32)
    for s in range(256):
        if hist[s]:
            d, i = 0, leaf_id[s]
            while nsym > 1 and i != root:
                i = parent[i]; d += 1
            lens[s] = d if nsym > 1 else 1
    return lens


def _histograms():
    rng = np.random.default_rng(20261019)
    hs = [np.full(256, 1), np.full(256, 100), np.arange(1, 257), 2 ** (np.arange(256) % 20)]
    for hi in (1, 2, 3, 8, 100):
        hs += [rng.integers(0, hi + 1, 256), rng.integers(1, hi + 1, 256)]
    for m in (1, 2, 3, 5, 31, 32, 33, 64, 127, 128, 129, 255):
        h = np.zeros(256, dtype=np.int64)
        h[rng.choice(256, m, replace=False)] = rng.integers(1, 4, m)
        hs.append(h)
    p = np.arange(1, 257, dtype=np.float64) ** -1.48            # the benchmark's shape: Zipf, 16 MiB and 64 KiB chunks
    p /= p.sum()
    hs += [rng.multinomial(1 << 24, p[rng.permutation(256)]), rng.multinomial(1 << 16, p[rng.permutation(256)])]
    h = rng.integers(0, 3, 256)
    h[rng.choice(256, 4, replace=False)] = (1 << 29) - 200      # counts near 2^31 in total: 64-bit keys
    hs.append(h)
    return [np.asarray(h, dtype=np.int64) for h in hs]


@pytest.mark.parametrize("idx", range(len(_histograms())))
def test_warp_replay_model_matches_the_oracle(idx):
    h = _histograms()[idx]
    want, _ = orc.code_lengths(h.astype(np.uint64))
    assert np.array_equal(replay_lengths(h), want)
