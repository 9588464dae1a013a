"""GPU parity tests: every stage of the CUDA hot path (through the C ABI of libhuffb200.so)
against the CPU oracle, bit-exact.  Mirrors the reference's tests."""
import hashlib
import json
import os

import numpy as np
import pytest

import datasets
import orc

pytestmark = pytest.mark.gpu
MiB = 1 << 20


def oracle_encode(data, chunk):
    """Per-chunk oracle encode -> (payload, comp_off[K+1], lens[K,256])."""
    data = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data
    K = (data.size + chunk - 1) // chunk
    parts, lens, off = [], np.zeros((K, 256), dtype=np.uint8), np.zeros(K + 1, dtype=np.uint64)
    for k in range(K):
        p, ln, _ = orc.encode_chunk(data[k * chunk:(k + 1) * chunk])
        parts.append(p); lens[k] = ln; off[k + 1] = off[k] + p.size
    return (np.concatenate(parts) if parts else np.zeros(0, np.uint8)), off, lens


def check_encode(codec, data, chunk):
    data = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data
    payload, off, lens, hist = codec.encode(data, chunk, want_hist=True)
    rp, roff, rlens = oracle_encode(data, chunk)
    K = len(roff) - 1
    for k in range(K):
        assert np.array_equal(hist[k].astype(np.uint64), orc.histogram(data[k * chunk:(k + 1) * chunk])), "hist chunk %d" % k
    assert np.array_equal(lens, rlens), "code lengths"
    assert np.array_equal(off, roff), "chunk offsets"
    assert np.array_equal(payload, rp), "payload bytes"
    # and back
    sizes = np.diff(off).astype(np.uint32)
    orig = np.array([min(chunk, data.size - k * chunk) for k in range(K)], dtype=np.uint32)
    back = codec.decode(payload, off[:-1], sizes, orig, lens)
    assert np.array_equal(back, data), "round trip"
    return payload, off, lens


# ---- histogram (CpuFrequencyServiceTest.java) ------------------------------------------------------
def test_histogram_known_answers(hz, codec):
    fs = hz.B200FrequencyService(codec)
    h = fs.compute_histogram(bytes([0, 1, 2, 1, 0, 1]), 0, 6)
    assert (h[0], h[1], h[2]) == (2, 3, 1) and h.sum() == 6 and h.dtype == np.int64
    assert np.all(fs.compute_histogram(bytes(range(256)), 0, 256) == 1)
    d = bytes(5 if i < 50 else 10 for i in range(100))
    h = fs.compute_histogram(d, 25, 50)
    assert h[5] == 25 and h[10] == 25
    h = fs.compute_histogram(np.array([-1, -2, -3, -1], dtype=np.int8).view(np.uint8), 0, 4)
    assert (h[255], h[254], h[253]) == (2, 1, 1)
    d = (np.arange(128 * 1024) % 10).astype(np.uint8)
    assert np.array_equal(fs.compute_histogram(d, 0, d.size).astype(np.uint64), orc.histogram(d))
    assert fs.is_available() and "B200" in fs.get_service_name()


@pytest.mark.parametrize("n,chunk", [(1, 1), (15, 64), (61440, 61440), (61441, 61440), (57344, 57344), (57345, 57344), (200_001, 50_000),
                                      (3 * MiB + 17, MiB), (1_000_003, 333_337)])
@pytest.mark.parametrize("kind", ["uniform", "zipf", "same"])
def test_histogram_exact(codec, n, chunk, kind):
    rng = np.random.default_rng(n + chunk)
    data = {"uniform": lambda: rng.integers(0, 256, n, dtype=np.uint8),
            "zipf": lambda: datasets.zipf_stream(n, 3, seed=n),
            "same": lambda: np.full(n, 0x41, dtype=np.uint8)}[kind]()
    hist = codec.histogram(data, chunk)
    for k in range(hist.shape[0]):
        assert np.array_equal(hist[k].astype(np.uint64), orc.histogram(data[k * chunk:(k + 1) * chunk]))


def test_histogram_unaligned_device_pointer(codec):
    import torch
    rng = np.random.default_rng(3)
    host = rng.integers(0, 256, 500_000, dtype=np.uint8)
    dev = torch.from_numpy(host).cuda()
    for shift in (1, 7, 13):
        view = dev[shift:]
        hist = torch.zeros((4, 256), dtype=torch.int32, device="cuda")
        n = view.numel()
        chunk = (n + 3) // 4
        codec.histogram_raw(view.data_ptr(), n, chunk, hist)
        codec.sync()
        h = hist.cpu().numpy().astype(np.uint64)
        for k in range(4):
            assert np.array_equal(h[k], orc.histogram(host[shift:][k * chunk:(k + 1) * chunk]))


# ---- codebook (CanonicalHuffmanTest.java, HuffmanPropertyTest.java) ----------------------------------
def _check_codebooks(codec, hists):
    hists = np.asarray(hists, dtype=np.uint32).reshape(-1, 256)
    lens, codes = codec.build_codebooks(hists)
    for k in range(hists.shape[0]):
        ln, mx = orc.code_lengths(hists[k].astype(np.uint64))
        cd, _ = orc.canonical_codes(ln)
        assert mx >= 0
        assert np.array_equal(lens[k], ln.astype(np.uint8)), "lengths, histogram %d" % k
        assert np.array_equal(codes[k], cd), "codes, histogram %d" % k
    assert np.array_equal(codec.codes_from_lengths(lens), codes)


def test_codebook_jqwik_style(codec):
    rng = np.random.default_rng(2024)
    h = rng.integers(0, 1001, (400, 256)).astype(np.uint32)          # HuffmanPropertyTest.java:81-92
    h[np.arange(400), rng.integers(0, 256, 400)] += 1
    _check_codebooks(codec, h)


def test_codebook_heavy_ties(codec):
    rng = np.random.default_rng(7)
    hs = [np.full(256, 100), np.full(256, 1), 2 ** (np.arange(256) % 20), rng.choice([3, 5], 256),
          rng.choice([1, 2, 4, 8], 256), np.arange(1, 257), np.arange(256, 0, -1),
          np.where(np.arange(256) % 3 == 0, 0, 7), rng.integers(0, 4, 256)]
    for nsym in (2, 3, 4, 5, 7, 8, 9, 31, 32, 33, 100, 255):
        h = np.zeros(256, dtype=np.int64); h[rng.choice(256, nsym, replace=False)] = 1; hs.append(h)
        h = np.zeros(256, dtype=np.int64); h[:nsym] = rng.integers(1, 4, nsym); hs.append(h)
    _check_codebooks(codec, np.array(hs, dtype=np.uint32))


def test_codebook_edge_cases(codec):
    hs = np.zeros((4, 256), dtype=np.uint32)
    hs[1, 42] = 1000                                                 # single symbol -> length 1, code 0
    hs[2, 0], hs[2, 255] = 1, 1
    hs[3, 10], hs[3, 20], hs[3, 30] = 5, 5, 10
    lens, codes = codec.build_codebooks(hs)
    assert not lens[0].any() and not codes[0].any()
    assert lens[1, 42] == 1 and lens[1].sum() == 1 and codes[1, 42] == 0
    _check_codebooks(codec, hs)


def test_codebook_long_codes(codec, hz):
    hs = [datasets.fib_like_hist(n).astype(np.uint32) for n in (10, 17, 24, 33)]     # max length 9..32
    _check_codebooks(codec, np.array(hs))
    assert orc.code_lengths(datasets.fib_like_hist(33))[1] == 32
    with pytest.raises(hz.HzError) as e:                                             # depth 33: reference throws
        codec.build_codebooks(datasets.fib_like_hist(34).astype(np.uint32))
    assert e.value.status == hz.HZ_ERR_CODE_TOO_LONG


# ---- encode + decode ---------------------------------------------------------------------------------
@pytest.mark.parametrize("idx", range(11))
def test_reference_inputs(codec, decmode, idx):
    data, name, chunk, _ = datasets.reference_cases()[idx]
    check_encode(codec, data, chunk)
    check_encode(codec, data, 1 * MiB)


def test_fixtures(codec, decmode):
    for chunk in (1 * MiB, 2 * MiB, 16 * MiB, 32 * MiB):
        payload, off, lens = check_encode(codec, datasets.fixture_bytes("test_2mb.bin"), chunk)
        assert not payload.any() and payload.size == 262144
    payload, _, lens = check_encode(codec, datasets.fixture_bytes("test_small.bin"), 32 * MiB)
    assert payload.size == 256 and lens[0, 65] == 1
    # test_input.bin, the reference's own 1 MiB fixture: all 256 code lengths are 8, so code[s] == s and the payload
    # IS the input; .dcz = payload + 68 + len(name) + 572 + 8 bytes
    uni = np.frombuffer(datasets.fixture_bytes("test_input.bin"), dtype=np.uint8)
    for chunk in (16 * MiB, 32 * MiB):
        payload, _, lens = check_encode(codec, uni, chunk)
        assert (lens[0] == 8).all() and np.array_equal(payload, uni)
    z = codec.compress_buffer(uni, 16 * MiB, "test_input.bin", 0)
    assert len(z) == 1048576 + 68 + len("test_input.bin") + 572 + 8 and z == orc.compress(uni, 16 * MiB, "test_input.bin", 0)
    assert codec.decompress_buffer(z) == uni.tobytes()
    check_encode(codec, orc.java_random_bytes(7, 1 * MiB), 16 * MiB)


@pytest.mark.parametrize("entropy", [1, 2, 3, 4, 5, 6, 7, 8])
@pytest.mark.parametrize("chunk", [64 * 1024, 1 * MiB])
def test_zipf_streams(codec, decmode, entropy, chunk):
    data = datasets.zipf_stream(2 * MiB + 4321, entropy, seed=entropy)
    check_encode(codec, data, chunk)


@pytest.mark.parametrize("n,chunk", [(0, 1024), (1, 1024), (7, 3), (1023, 4096), (61439, 1 << 20), (61440, 1 << 20),
                                      (61441, 1 << 20), (122880, 61440), (57343, 1 << 20), (57344, 1 << 20), (57345, 1 << 20), (114688, 57344), (8191, 8192), (8193, 1 << 20), (1_000_003, 100_003), (300_000, 77)])
def test_ragged_sizes(codec, decmode, n, chunk):
    data = datasets.zipf_stream(n, 4, seed=n + 1) if n else np.zeros(0, np.uint8)
    check_encode(codec, data, chunk)


def test_wide_codes_encode_decode(codec, decmode):
    """Chunks whose longest code is > 16 bits (wide encoder path) and > 12 bits (decoder fallback)."""
    rng = np.random.default_rng(11)
    for nsym in (18, 24, 28, 29, 30):       # longest code 17, 23, 27 (medium path) and 28, 29 (wide path)
        f = datasets.fib_like_hist(nsym)[:nsym].astype(np.int64)      # exact Fibonacci counts -> depth nsym-1
        data = np.repeat(np.arange(nsym, dtype=np.uint8), f)
        rng.shuffle(data)
        payload, off, lens = check_encode(codec, data, 1 << 22)
        assert lens.max() == nsym - 1


def test_uniform_length_codes(codec):
    """Equal-length codes never self-synchronise; lengths 7 / 5 / 3 do not divide the subsequence size."""
    for nsym in (128, 32, 8, 2):
        data = (np.arange(700_001) * 2654435761 % nsym).astype(np.uint8)
        payload, off, lens = check_encode(codec, data, 1 << 20)
        assert len(set(lens[0][lens[0] > 0])) == 1


def test_decode_oracle_streams_and_errors(codec, decmode, hz):
    data = datasets.zipf_stream(500_000, 5, seed=9)
    rp, roff, rlens = oracle_encode(data, 200_000)
    sizes = np.diff(roff).astype(np.uint32)
    orig = np.array([200_000, 200_000, 100_000], dtype=np.uint32)
    assert np.array_equal(codec.decode(rp, roff[:-1], sizes, orig, rlens), data)
    # single-symbol chunk with a stray 1 bit: "Huffman decode error at position i"
    ln = np.zeros((1, 256), dtype=np.uint8); ln[0, 65] = 1
    comp = np.zeros(64, dtype=np.uint8); comp[10] = 0x10
    with pytest.raises(hz.HzError) as e:
        codec.decode(comp, [0], [64], [512], ln)
    assert e.value.status == hz.HZ_ERR_DECODE
    out = codec.decode(np.zeros(64, np.uint8), [0], [64], [512], ln)      # context still usable
    assert np.all(out == 65)
    # over-subscribed length table
    bad = np.zeros((1, 256), dtype=np.uint8); bad[0, :3] = 1
    with pytest.raises(hz.HzError) as e:
        codec.decode(comp, [0], [64], [8], bad)
    assert e.value.status == hz.HZ_ERR_BAD_LENGTHS


@pytest.mark.parametrize("H,chunk", [(5, 1 * MiB), (2, 300_000), (7, 70_000)])
def test_decode_damaged_streams_match_the_oracle(codec, decmode, H, chunk):
    """A complete prefix code decodes ANY bit string, so a damaged stream has a well-defined reference result:
    flipped bits (the self-synchronisation guesses fail and are repaired inside CTAs and across CTA boundaries),
    a truncated payload (zero bits past the end) and a short orig_size must give the oracle's bytes."""
    rng = np.random.default_rng(H * 1000 + 7)
    data = datasets.zipf_stream(3 * chunk, H, seed=H + 60)
    payload, off, lens = codec.encode(data, chunk)[:3]
    sizes = np.diff(off).astype(np.uint32)
    for k in range(3):
        comp = payload[int(off[k]):int(off[k + 1])].copy()
        ln = lens[k].astype(np.int32)
        for trial in range(3):
            bad = comp.copy()
            flips = rng.integers(0, bad.size * 8, 25)
            for f in flips:
                bad[f >> 3] ^= 0x80 >> (f & 7)
            cut = int(rng.integers(0, 4))                      # drop up to 3 bytes at the end
            bad = bad[:bad.size - cut]
            n_out = chunk - int(rng.integers(0, 3)) * 1000      # sometimes fewer symbols than the stream holds
            ref, rc = orc.decode(bad, ln, n_out, literal=False)
            assert rc == 0
            out = codec.decode(bad, [0], [bad.size], [n_out], lens[k:k + 1])
            assert np.array_equal(out, ref), "chunk %d trial %d" % (k, trial)


@pytest.mark.parametrize("warps", ["24", "8", "5", "24 as cluster pairs"])
def test_fused_decoder_cta_shapes(codec, knob, warps):
    """The fused decoder picks its CTA shape by the units per chunk (24 warps x 1 CTA per SM, 8 x 2, 5 x 3;
    HZ_FU_WARPS forces one; with fewer chunks than SMs the 24-warp grid is launched as cluster pairs, HZ_FU_CLUSTER=1).  Every shape on: thousands of small chunks of mixed entropy with a ragged tail, chunks
    whose payload ends exactly on / one byte around a subsequence and a unit boundary, a payload that holds MORE
    symbols than orig_size and one that holds FEWER (the chunk's last subsequence stops at the chunk's end and the
    all-zero codeword's symbol fills the rest, TableBasedHuffmanDecoder.java:204-208), damaged streams."""
    knob("HZ_DEC", "fused")
    knob("HZ_FU_WARPS", warps.split()[0])
    if "pairs" in warps:                                   # the launch of streams with fewer chunks than SMs (DSMEM ring pushes)
        knob("HZ_FU_CLUSTER", "1")
    rng = np.random.default_rng(int(warps.split()[0]) + 40)
    # 1. many small chunks
    parts = [datasets.zipf_stream(200_000, H, seed=70 + H) for H in (1, 3, 4, 6, 7)]
    data = np.concatenate(parts + [datasets.zipf_stream(12_345, 5, seed=80)])
    for chunk in (3_000, 16 * 1024, 65_536):
        check_encode(codec, data, chunk)
    # 2. payload sizes around subsequence (S words) and unit (32 S words) boundaries: trim the symbol count until the
    #    chunk's compressed size hits the wanted value, then decode against the oracle with longer / shorter orig_size
    base = datasets.zipf_stream(400_000, 4, seed=91)
    for target in (17 * 4 * 32 * 9, 17 * 4 * 32 * 9 + 1, 17 * 4 * 32 * 9 - 1, 17 * 4 * 40, 17 * 4 * 40 + 3):
        n = int(target * 8 / 4.03)
        for _ in range(60):
            comp, ln, _ = orc.encode_chunk(base[:n])
            if comp.size == target: break
            n += int((target - comp.size) * 8 / 4.03) or (1 if comp.size < target else -1)
        lens = ln.astype(np.uint8)[None, :]
        for n_out in (n, n - 7, n + 5000):
            ref, rc = orc.decode(comp, ln.astype(np.int32), n_out, literal=False)
            assert rc == 0
            out = codec.decode(comp, [0], [comp.size], [n_out], lens)
            assert np.array_equal(out, ref), "payload %d bytes, %d of %d symbols" % (comp.size, n_out, n)
    # 3. damaged small chunks
    chunk = 50_000
    d3 = datasets.zipf_stream(8 * chunk, 5, seed=93)
    payload, off, lens = codec.encode(d3, chunk)[:3]
    bad = payload.copy()
    for f in rng.integers(0, bad.size * 8, 200):
        bad[f >> 3] ^= 0x80 >> (f & 7)
    sizes = np.diff(off).astype(np.uint32)
    out = codec.decode(bad, off[:-1], sizes, [chunk] * 8, lens)
    for k in range(8):
        ref, rc = orc.decode(bad[int(off[k]):int(off[k + 1])], lens[k].astype(np.int32), chunk, literal=False)
        assert rc == 0 and np.array_equal(out[k * chunk:(k + 1) * chunk], ref), "damaged chunk %d" % k


def test_decode_reads_zero_bits_past_the_end(codec, decmode):
    # TableBasedHuffmanDecoder.java:204-208: bits past the end of the chunk are 0
    ln = np.zeros((1, 256), dtype=np.uint8); ln[0, 7] = 1; ln[0, 9] = 1      # 7 -> '0', 9 -> '1'
    comp = np.array([0b10100000], dtype=np.uint8)
    out = codec.decode(comp, [0], [1], [20], ln)
    ref, rc = orc.decode(comp, ln[0].astype(np.int32), 20, literal=True)
    assert rc == 0 and np.array_equal(out, ref)


def _perm_blocks(rng, nbytes):
    """Bytes in which every value occurs equally often per 256-byte block (-> all 256 code lengths are 8)."""
    nb = (nbytes + 255) // 256
    return np.concatenate([rng.permutation(256).astype(np.uint8) for _ in range(nb)])[:nbytes]


@pytest.mark.parametrize("chunk", [4096, 65536 + 256, 1 * MiB + 512])
def test_identity_chunks_take_the_copy_path(codec, decmode, chunk):
    """A chunk whose 256 symbols all get 8-bit codes has code[s] == s: encode and decode are byte copies
    (hz_group_copy).  Mixed with ordinary chunks so that chunk offsets are odd (every source / destination
    misalignment), with ragged tails, against the oracle bit for bit."""
    rng = np.random.default_rng(chunk)
    K = 9
    parts = []
    for k in range(K):
        if k % 3 == 1:
            parts.append(datasets.zipf_stream(chunk, 1 + k % 7, seed=k))          # ordinary chunk, odd payload size
        else:
            parts.append(_perm_blocks(rng, chunk))
    parts.append(_perm_blocks(rng, 256 * 5))                                      # short identity tail chunk
    data = np.concatenate(parts)
    payload, off, lens = check_encode(codec, data, chunk)
    ident = [k for k in range(K + 1) if (lens[k] == 8).all()]
    assert len(ident) >= 6
    for k in ident:
        assert np.array_equal(payload[int(off[k]):int(off[k + 1])], data[k * chunk:(k + 1) * chunk])
    # a truncated identity chunk: missing bytes decode as symbol 0 (zero bits past the end), like the oracle
    k = ident[1]
    comp = payload[int(off[k]):int(off[k + 1])].copy()
    n = comp.size
    out = codec.decode(comp[:n - 100], [0], [n - 100], [n], lens[k:k + 1])
    ref, rc = orc.decode(comp[:n - 100], lens[k].astype(np.int32), n, literal=True)
    assert rc == 0 and np.array_equal(out, ref)
    assert np.array_equal(out[:n - 100], comp[:n - 100]) and not out[n - 100:].any()


def test_deterministic(codec):
    data = datasets.zipf_stream(3 * MiB, 4, seed=77)
    a = codec.encode(data, MiB)
    for _ in range(3):
        b = codec.encode(data, MiB)
        assert all(np.array_equal(x, y) for x, y in zip(a, b))


def test_global_codebook_extension(codec):
    data = datasets.zipf_stream(1_000_000, 4, seed=5)
    hist = codec.histogram(data, 250_000).astype(np.uint64).sum(axis=0)
    lens, codes = codec.build_codebooks(hist.astype(np.uint32))
    payload, off = codec.encode_with_lengths(data, 250_000, lens[0])
    ln = lens[0].astype(np.int32)
    cd, _ = orc.canonical_codes(ln)
    for k in range(4):
        ref = orc.encode(data[k * 250_000:(k + 1) * 250_000], ln, cd)
        assert np.array_equal(payload[int(off[k]):int(off[k + 1])], ref)
    back = codec.decode(payload, off[:-1], np.diff(off).astype(np.uint32), np.full(4, 250_000, np.uint32),
                        np.tile(lens[0], (4, 1)))
    assert np.array_equal(back, data)


def test_stage_metrics_of_the_file_calls(codec, hz):
    """getLastStageMetrics() (cpu/CpuCompressionService.java:52): after compress the stages of
    model/StageMetrics.java:11-20 that the reference records there carry time, after decompress those it records
    there; names and order are the enum's."""
    data = datasets.zipf_stream(5 * MiB + 3, 4, seed=21)
    z = codec.compress_buffer(data, MiB, "m.bin", 0)
    m = codec.stage_metrics()
    assert list(m) == ["FREQUENCY_ANALYSIS", "HUFFMAN_TREE_BUILD", "ENCODING", "CHECKSUM_COMPUTE", "FILE_IO", "HEADER_WRITE",
                       "DECODING", "CHECKSUM_VERIFY"]
    for st in ("FREQUENCY_ANALYSIS", "HUFFMAN_TREE_BUILD", "ENCODING", "CHECKSUM_COMPUTE", "FILE_IO", "HEADER_WRITE"):
        assert m[st][0] > 0 and m[st][1] > 0, st
    assert m["DECODING"][1] == 0 and m["ENCODING"][2] == data.size and m["CHECKSUM_COMPUTE"][2] == data.size
    assert codec.decompress_buffer(z) == data.tobytes()
    m = codec.stage_metrics()
    for st in ("HUFFMAN_TREE_BUILD", "DECODING", "CHECKSUM_VERIFY", "FILE_IO"):
        assert m[st][0] > 0 and m[st][1] > 0, st
    assert m["ENCODING"][1] == 0 and m["DECODING"][2] == data.size
    svc = hz.B200CompressionService.__new__(hz.B200CompressionService)
    svc._codec = codec
    assert svc.get_last_stage_metrics() == m


def test_encode_global_in_library(codec):
    """hz_encode_global (one histogram pass, device-side reduction, codebook and encode in one call; the NCCL
    all-reduce is skipped without a communicator) == histogram + hz_build_codebooks + hz_encode_with_lengths, and
    the lengths are the oracle's for the summed histogram.  An expanding code (a rare shard coded with a table
    built elsewhere) and an empty shard work too."""
    data = datasets.zipf_stream(3_000_001, 3, seed=15)
    chunk = 400_000
    payload, off, l256 = codec.encode_global(data, chunk)
    hist = codec.histogram(data, chunk).astype(np.uint64).sum(axis=0)
    ref_len = orc.code_lengths(hist)[0].astype(np.uint8)
    assert np.array_equal(l256, ref_len)
    p2, o2 = codec.encode_with_lengths(data, chunk, l256)
    assert np.array_equal(off, o2) and np.array_equal(payload, p2)
    K = len(off) - 1
    # (8 chunks: both calls take the chained sizes kernel + dependent encoder) every chunk against the oracle's bit packer
    cd, _ = orc.canonical_codes(l256.astype(np.int32))
    for k in range(K):
        ref = orc.encode(data[k * chunk:(k + 1) * chunk], l256.astype(np.int32), cd)
        assert np.array_equal(payload[int(off[k]):int(off[k + 1])], ref), k
    orig = np.array([min(chunk, data.size - k * chunk) for k in range(K)], dtype=np.uint32)
    back = codec.decode(payload, off[:-1], np.diff(off).astype(np.uint32), orig, np.tile(l256, (K, 1)))
    assert np.array_equal(back, data)
    p0, o0, l0 = codec.encode_global(np.zeros(0, np.uint8), chunk)
    assert p0.size == 0 and o0.tolist() == [0] and not l0.any()


# ---- container (.dcz) ------------------------------------------------------------------------------------
@pytest.mark.parametrize("idx", range(11))
def test_dcz_byte_identical(codec, idx):
    data, name, chunk, expect = datasets.reference_cases()[idx]
    z = codec.compress_buffer(data, chunk, name, 0)
    assert len(z) == expect
    assert z == orc.compress(data, chunk, name, 0)
    golden = json.load(open(os.path.join(datasets.GOLDEN_DIR, "reference_cases.json")))[idx]
    assert hashlib.sha256(z).hexdigest() == golden["dcz_sha256"]
    assert codec.decompress_buffer(z) == data
    assert orc.decompress(z) == data


def test_service_files(hz, tmp_path):
    data = datasets.zipf_stream(5 * MiB + 99, 4, seed=21).tobytes()
    src = tmp_path / "in.bin"; src.write_bytes(data)
    with hz.B200CompressionService(1) as svc:
        seen = []
        svc.compress(str(src), str(tmp_path / "out.dcz"), lambda f: seen.append(f))
        assert len(seen) == 6 and seen[-1] == 1.0 and all(0 < f <= 1 for f in seen)
        z = (tmp_path / "out.dcz").read_bytes()
        mtime = int.from_bytes(z[int.from_bytes(z[-8:], "big") + 12 + 6 + 8:][:8], "big")
        assert z == orc.compress(data, MiB, "in.bin", mtime)
        assert svc.verify_integrity(str(tmp_path / "out.dcz"))
        svc.decompress(str(tmp_path / "out.dcz"), str(tmp_path / "back.bin"))
        assert (tmp_path / "back.bin").read_bytes() == data
        with pytest.raises(NotImplementedError):
            svc.resume_compression(str(src), str(tmp_path / "o2"), 0)
        # corrupt one payload byte -> checksum / decode failure, verify says False
        bad = bytearray(z); bad[1000] ^= 0x5A
        (tmp_path / "bad.dcz").write_bytes(bytes(bad))
        assert not svc.verify_integrity(str(tmp_path / "bad.dcz"))
        with pytest.raises(hz.HzError):
            svc.decompress(str(tmp_path / "bad.dcz"), str(tmp_path / "bad.out"))
        # bad magic
        (tmp_path / "junk.dcz").write_bytes(b"not a dcz file at all" * 10)
        with pytest.raises(hz.HzError) as e:
            svc.decompress(str(tmp_path / "junk.dcz"), str(tmp_path / "junk.out"))
        assert e.value.status == hz.HZ_ERR_FORMAT


@pytest.mark.parametrize("n,chunk_mib", [(300 * MiB + 7, 16), (260 * MiB, 1)])
def test_service_files_multi_batch(hz, tmp_path, n, chunk_mib):
    """Files larger than one 128 MiB batch go through the double-buffered file pipeline (reader task, GPU +
    SHA-256, writer task over two pinned slots): the .dcz must still be byte-identical to the oracle's."""
    data = datasets.zipf_stream(n, 4, seed=77).tobytes()
    src = tmp_path / "big.bin"; src.write_bytes(data)
    with hz.B200CompressionService(chunk_mib) as svc:
        svc.compress(str(src), str(tmp_path / "big.dcz"), None)
        z = (tmp_path / "big.dcz").read_bytes()
        mtime = int.from_bytes(z[int.from_bytes(z[-8:], "big") + 12 + 7 + 8:][:8], "big")
        assert z == orc.compress(data, chunk_mib * MiB, "big.bin", mtime)
        assert svc.verify_integrity(str(tmp_path / "big.dcz"))
        svc.decompress(str(tmp_path / "big.dcz"), str(tmp_path / "back.bin"))
        assert (tmp_path / "back.bin").read_bytes() == data
        bad = bytearray(z); bad[200 * MiB // 2] ^= 0x5A              # a payload byte of a later batch
        (tmp_path / "bad.dcz").write_bytes(bytes(bad))
        assert not svc.verify_integrity(str(tmp_path / "bad.dcz"))


def test_container_thousands_of_small_chunks_hash_on_the_gpu(codec, hz):
    """>= 1536 chunks per batch: the per-chunk SHA-256 comes from the GPU kernel (compress: on the batch already in
    device memory; decompress / verify: on the decoded batch before it leaves the device).  The container must still
    be the oracle's byte for byte, a flipped payload byte must still be caught, and the ragged last chunk counts."""
    data = datasets.zipf_stream(40 * MiB + 1234, 5, seed=31).tobytes()
    chunk = 16 * 1024
    z = codec.compress_buffer(data, chunk, "small.bin", 7)
    assert z == orc.compress(data, chunk, "small.bin", 7)
    m = codec.stage_metrics()
    assert m["CHECKSUM_COMPUTE"][0] > 0
    assert codec.decompress_buffer(z) == data
    assert codec.stage_metrics()["CHECKSUM_VERIFY"][0] > 0
    bad = bytearray(z); bad[12345] ^= 1
    with pytest.raises(hz.HzError) as e:
        codec.decompress_buffer(bytes(bad))
    assert e.value.status in (hz.HZ_ERR_CHECKSUM, hz.HZ_ERR_DECODE)


def test_untrusted_metadata_is_rejected_not_followed(codec, decmode, hz):
    """Chunk offsets / sizes reach hz_decode from a footer nobody vouches for: a chunk that does not lie inside the
    stream is reported (HZ_ERR_ARG) instead of being read; the context stays usable.  Same for a footer whose
    compressedOffset wraps around 2^64 (buffer API)."""
    data = datasets.zipf_stream(600_000, 4, seed=8)
    payload, off, lens = codec.encode(data, 200_000)[:3]
    sizes = np.diff(off).astype(np.uint32)
    orig = np.full(3, 200_000, np.uint32)
    for bad_off, bad_size in (([0, int(off[1]), payload.size - 10], sizes), (off[:-1], [sizes[0], sizes[1], sizes[2] + 4096]),
                              ([0, 2**63, int(off[2])], sizes)):
        with pytest.raises(hz.HzError) as e:
            codec.decode(payload, np.array(bad_off, dtype=np.uint64), np.array(bad_size, dtype=np.uint32), orig, lens)
        assert e.value.status == hz.HZ_ERR_ARG
    assert np.array_equal(codec.decode(payload, off[:-1], sizes, orig, lens), data)
    z = bytearray(orc.compress(data.tobytes(), 200_000, "x.bin", 0))
    fo = int.from_bytes(z[-8:], "big")
    rec0 = fo + 4 + 4 + 4 + 5 + 8 + 8 + 4 + 32 + 4            # first chunk record (name "x.bin")
    z[rec0 + 16: rec0 + 24] = (2**64 - 8).to_bytes(8, "big")   # compressedOffset of chunk 0
    with pytest.raises(hz.HzError):
        codec.decompress_buffer(bytes(z))


def test_fixed_length_encode_of_a_large_host_buffer_may_expand(codec):
    """hz_encode_with_lengths on >= 128 MiB of HOST data takes the pipelined path; a table built for other data can
    expand a batch beyond its input size (here 9/8): the slots must hold that, and the result must equal the
    device-resident path's."""
    import torch
    n, chunk = 160 * MiB, 4 * MiB
    lens256 = np.full(256, 9, dtype=np.uint8)        # 256 x 9 bits (Kraft sum 0.5): every byte costs 9 bits
    data = datasets.zipf_stream(n, 8, seed=4)
    payload, off = codec.encode_with_lengths(data, chunk, lens256)                   # host path (pipelined)
    assert int(off[-1]) == payload.size > n
    d = torch.from_numpy(data).cuda()
    K = len(off) - 1
    out = torch.empty(2 * n, dtype=torch.uint8, device="cuda")
    doff = torch.zeros(K + 1, dtype=torch.int64, device="cuda")
    codec.encode_with_lengths_raw(d.data_ptr(), n, chunk, torch.from_numpy(lens256).cuda().data_ptr(), out.data_ptr(), 2 * n, doff.data_ptr())
    codec.sync()
    assert np.array_equal(doff.cpu().numpy().astype(np.uint64), off)
    assert np.array_equal(out[: int(off[-1])].cpu().numpy(), payload)


def test_legacy_header_first_layout(codec):
    data = b"Hello World! " * 100
    z = orc.compress(data, MiB, "t.txt", 5)
    payload_len = int.from_bytes(z[-8:], "big")
    legacy = z[payload_len:-8] + z[:payload_len]                 # header first, payload after
    assert codec.decompress_buffer(legacy) == data
    assert orc.decompress(legacy) == data


def test_sha256_chunks(codec):
    data = datasets.zipf_stream(300_000, 6, seed=3)
    dig = codec.sha256_chunks(data, 64 * 1024)
    for k in range(dig.shape[0]):
        assert dig[k].tobytes() == hashlib.sha256(data[k * 65536:(k + 1) * 65536].tobytes()).digest()


# ---- large, device-resident (size-independent properties) --------------------------------------------------
def test_large_device_resident_roundtrip(codec):
    import torch
    n, chunk = 256 * MiB, 16 * MiB
    K = n // chunk
    qt = datasets.zipf_qtable(4)
    src = torch.empty(n, dtype=torch.uint8, device="cuda")
    codec.synth_fill(src.data_ptr(), n, 0, 0x5EED0001, qt)
    codec.sync()
    assert np.array_equal(src[:100_000].cpu().numpy(), datasets.synth_host(100_000, 0x5EED0001, qt))
    comp = torch.empty(n + 16, dtype=torch.uint8, device="cuda")
    off = torch.zeros(K + 1, dtype=torch.int64, device="cuda")
    lens = torch.zeros((K, 256), dtype=torch.uint8, device="cuda")
    codec.encode_raw(src.data_ptr(), n, chunk, comp.data_ptr(), n, off.data_ptr(), lens.data_ptr(), None)
    codec.sync()
    offh = off.cpu().numpy().astype(np.uint64)
    # chunk 0 and the last chunk against the oracle, all chunks through the round trip
    for k in (0, K - 1):
        ref, ln, _ = orc.encode_chunk(src[k * chunk:(k + 1) * chunk].cpu().numpy())
        assert np.array_equal(lens[k].cpu().numpy(), ln.astype(np.uint8))
        assert np.array_equal(comp[int(offh[k]):int(offh[k + 1])].cpu().numpy(), ref)
    sizes = (off[1:] - off[:-1]).to(torch.int32)
    orig = torch.full((K,), chunk, dtype=torch.int32, device="cuda")
    back = torch.zeros(n, dtype=torch.uint8, device="cuda")
    codec.decode_raw(comp.data_ptr(), int(offh[K]), off.data_ptr(), sizes.data_ptr(), orig.data_ptr(), None,
                     lens.data_ptr(), K, back.data_ptr(), n)
    codec.sync()
    assert torch.equal(back, src)


def test_chained_encode_equals_separate_launches(codec, knob, hz):
    """Streams of >= 8 chunks of >= 2 MiB take the chained histogram -> codebook -> offsets kernel and an encoder that
    waits per chunk (programmatic stream serialization).  Its payload, offsets, lengths and chunk histograms must equal
    those of the separate launches (HZ_ENC_CHAIN=0) bit for bit - with a ragged last chunk, mixed entropies, an
    incompressible (identity) chunk and a one-symbol chunk, repeatedly (flags are reset per call) - and chunk 0 / the
    last chunk must equal the oracle's; a too small output buffer is reported, not overrun."""
    import torch
    chunk = 8 * MiB
    n = 9 * chunk + 12_345
    K = 10
    src = torch.empty(n, dtype=torch.uint8, device="cuda")
    for i, H in enumerate((4, 2, 8, 6, 1, 4, 7, 3, 5)):
        codec.synth_fill(src.data_ptr() + i * chunk, chunk, i * chunk, 0x5EED0100 + i, datasets.zipf_qtable(H))
    codec.synth_fill(src.data_ptr() + 9 * chunk, 12_345, 0, 0x5EED0200, datasets.zipf_qtable(4))
    src[4 * chunk:5 * chunk] = 0x41                        # a one-symbol chunk
    codec.sync()
    res = {}
    for mode in ("0", "1", "1"):
        knob("HZ_ENC_CHAIN", mode)
        comp = torch.zeros(n + 16, dtype=torch.uint8, device="cuda")
        off = torch.zeros(K + 1, dtype=torch.int64, device="cuda")
        lens = torch.zeros((K, 256), dtype=torch.uint8, device="cuda")
        hist = torch.zeros((K, 256), dtype=torch.int32, device="cuda")
        codec.encode_raw(src.data_ptr(), n, chunk, comp.data_ptr(), n, off.data_ptr(), lens.data_ptr(), hist.data_ptr())
        codec.sync()
        cur = (comp.cpu().numpy(), off.cpu().numpy(), lens.cpu().numpy(), hist.cpu().numpy())
        if mode in res:
            for a, b in zip(res[mode], cur): assert np.array_equal(a, b)
        res[mode] = cur
    for a, b, what in zip(res["0"], res["1"], ("payload", "offsets", "lengths", "histograms")):
        assert np.array_equal(a, b), what
    comp, off, lens, _ = res["1"]
    for k in (0, 2, 4, K - 1):
        ref, ln, _ = orc.encode_chunk(src[k * chunk:min(n, (k + 1) * chunk)].cpu().numpy())
        assert np.array_equal(lens[k], ln.astype(np.uint8)) and np.array_equal(comp[int(off[k]):int(off[k + 1])], ref), k
    # output capacity one byte short of the payload
    total = int(off[K])
    small = torch.zeros(total + 16, dtype=torch.uint8, device="cuda")
    o2 = torch.zeros(K + 1, dtype=torch.int64, device="cuda")
    with pytest.raises(hz.HzError) as e:
        codec.encode_raw(src.data_ptr(), n, chunk, small.data_ptr(), total - 1, o2.data_ptr(), None, None)
        codec.sync()
    assert e.value.status == hz.HZ_ERR_OUT_TOO_SMALL
    assert int(small[total - 1:].sum()) == 0               # nothing at or past the capacity was written
    codec.encode_raw(src.data_ptr(), n, chunk, small.data_ptr(), total, o2.data_ptr(), None, None)   # the context stays usable
    codec.sync()
    assert np.array_equal(small[:total].cpu().numpy(), comp[:total])


def test_chained_encode_repeated_and_concurrent_contexts(hz, codec):
    """The chain's flags live in per-context scratch and are reset per call: back-to-back calls of different shapes on
    one context, and two contexts encoding at the same time on two streams, must keep producing the payload of the
    separate launches (taken once per shape with HZ_ENC_CHAIN=0)."""
    import torch
    chunk = 8 * MiB
    shapes = [(8 * chunk, 4), (11 * chunk + 77, 6), (9 * chunk - 1, 2)]
    srcs, refs = [], []
    os.environ["HZ_ENC_CHAIN"] = "0"; codec.reload_knobs()
    try:
        for i, (n, H) in enumerate(shapes):
            K = (n + chunk - 1) // chunk
            src = torch.empty(n, dtype=torch.uint8, device="cuda")
            codec.synth_fill(src.data_ptr(), n, 0, 0x5EED0300 + i, datasets.zipf_qtable(H))
            comp = torch.zeros(n + 16, dtype=torch.uint8, device="cuda")
            off = torch.zeros(K + 1, dtype=torch.int64, device="cuda")
            codec.encode_raw(src.data_ptr(), n, chunk, comp.data_ptr(), n, off.data_ptr(), None, None)
            codec.sync()
            srcs.append(src); refs.append((comp, off))
    finally:
        os.environ.pop("HZ_ENC_CHAIN", None); codec.reload_knobs()
    other = hz.Codec(0)
    s2 = torch.cuda.Stream()
    other.set_stream(s2.cuda_stream)
    try:
        outs = []
        for it in range(12):
            for c, i in ((codec, it % 3), (other, (it + 1) % 3)):
                n = shapes[i][0]; K = (n + chunk - 1) // chunk
                comp = torch.zeros(n + 16, dtype=torch.uint8, device="cuda")
                off = torch.zeros(K + 1, dtype=torch.int64, device="cuda")
                c.encode_raw(srcs[i].data_ptr(), n, chunk, comp.data_ptr(), n, off.data_ptr(), None, None)
                outs.append((i, comp, off))
        codec.sync(); other.sync()
        for i, comp, off in outs:
            assert torch.equal(off, refs[i][1]) and torch.equal(comp, refs[i][0]), "shape %d" % i
    finally:
        other.close()


def test_config3_1gib_32mib_chunks_every_chunk_against_the_oracle(codec):
    """BASELINE config 3 (1 GiB Zipf, ~4 bits/symbol) at the CLI's default chunk size of 32 MiB: code lengths,
    compressed sizes and the SHA-256 of EVERY chunk's payload equal the oracle's (chunk-parallel fast coder), and
    the stream decodes back."""
    import torch
    n, chunk = 1024 * MiB, 32 * MiB
    K = n // chunk
    src = torch.empty(n, dtype=torch.uint8, device="cuda")
    codec.synth_fill(src.data_ptr(), n, 0, 0x5EED0003, datasets.zipf_qtable(4))
    comp = torch.empty(n + 16, dtype=torch.uint8, device="cuda")
    off = torch.zeros(K + 1, dtype=torch.int64, device="cuda")
    lens = torch.zeros((K, 256), dtype=torch.uint8, device="cuda")
    codec.encode_raw(src.data_ptr(), n, chunk, comp.data_ptr(), n, off.data_ptr(), lens.data_ptr(), None)
    codec.sync()
    offh = off.cpu().numpy().astype(np.uint64)
    host = src.cpu().numpy()
    rcomp, rsizes, rlens, _ = orc.encode_chunks_mt(host, chunk, literal=False)
    assert np.array_equal(np.diff(offh).astype(np.uint64), np.asarray(rsizes, dtype=np.uint64))
    assert np.array_equal(lens.cpu().numpy(), np.asarray(rlens, dtype=np.uint8).reshape(K, 256))
    got = comp[: int(offh[K])].cpu().numpy()
    for k in range(K):                                   # the oracle's chunk k sits at k * (chunk + 8)
        a = hashlib.sha256(got[int(offh[k]):int(offh[k + 1])]).digest()
        b = hashlib.sha256(rcomp[k * (chunk + 8): k * (chunk + 8) + int(rsizes[k])]).digest()
        assert a == b, "payload of chunk %d" % k
    sizes = (off[1:] - off[:-1]).to(torch.int32)
    orig = torch.full((K,), chunk, dtype=torch.int32, device="cuda")
    back = torch.zeros(n, dtype=torch.uint8, device="cuda")
    codec.decode_raw(comp.data_ptr(), int(offh[K]), off.data_ptr(), sizes.data_ptr(), orig.data_ptr(), None,
                     lens.data_ptr(), K, back.data_ptr(), n)
    codec.sync()
    assert torch.equal(back, src)


# ---- multi-GPU host logic bound to the CUDA codec (world size 1 here; gloo x2 in test_parallel_gloo.py) ----
@pytest.mark.parametrize("n,chunk", [(1_000_003, 100_000), (3 * MiB + 5, MiB), (0, 4096)])
def test_sharded_compressor_cuda_binding(hz, codec, n, chunk):
    import importlib
    par = importlib.import_module("huffb200.parallel")
    data = datasets.zipf_stream(n, 4, seed=21) if n else np.zeros(0, np.uint8)
    sc = par.ShardedCompressor(codec)
    out = sc.compress(data, n, chunk, "shard.bin", 42)
    assert out == orc.compress(data, chunk, "shard.bin", 42), "sharded container differs from the reference container"
    glob = sc.compress(data, n, chunk, "shard.bin", 42, global_codebook=True)
    assert codec.decompress_buffer(glob) == data.tobytes()          # decoded (and SHA-verified) on the GPU
    assert orc.decompress(glob) == data.tobytes()                    # and by the reference decoder


def test_sharded_two_ranks_one_gpu_equals_reference(hz, codec):
    """Two chunk ranges coded one after the other on the same GPU and assembled on the host give the
    reference container (what two ranks on two GPUs produce; ranks never exchange payload data)."""
    import importlib
    par = importlib.import_module("huffb200.parallel")
    n, chunk = 2_500_000, 300_000
    data = datasets.zipf_stream(n, 5, seed=33)
    shards = []
    for r in range(2):
        lo, hi = par.byte_range(n, chunk, 2, r)
        payload, off, lens = codec.encode(data[lo:hi], chunk)
        dig = [bytes(d) for d in codec.sha256_chunks(data[lo:hi], chunk)]
        K = len(off) - 1
        shards.append((payload.tobytes(), np.diff(off).astype(np.uint32).tolist(),
                       [min(chunk, hi - lo - k * chunk) for k in range(K)], dig, lens))
    sizes = [s for p in shards for s in p[1]]
    footer = par.write_footer("two.bin", n, 7, chunk, sizes, [o for p in shards for o in p[2]],
                              [d for p in shards for d in p[3]], np.concatenate([p[4] for p in shards]),
                              sum(len(p[0]) for p in shards))
    assert b"".join(p[0] for p in shards) + footer == orc.compress(data, chunk, "two.bin", 7)


# ---- host-buffer pipeline (batches of chunks over three streams) --------------------------------
@pytest.mark.parametrize("n,chunk,H", [(160 * MiB + 12345, 4 * MiB, 4), (200 * MiB, 16 * MiB, 6), (130 * MiB + 1, 64 * 1024, 2),
                                       (140 * MiB + 3, 1 * MiB, 8)])
def test_pipelined_host_buffers_match_device_path(codec, n, chunk, H):
    """hz_encode / hz_decode with HOST buffers >= 128 MiB take the pipelined path; the result must be
    byte-identical to the device-resident single-shot path, and chunks must match the oracle."""
    import torch
    data = datasets.zipf_stream(n, H, seed=n & 0xFFFF)
    K = (n + chunk - 1) // chunk
    # pipelined: numpy (pageable host) buffers
    payload, off, lens, hist = codec.encode(data, chunk, want_hist=True)
    # single shot: device buffers
    d_in = torch.from_numpy(data).cuda()
    d_out = torch.empty(n + 16, dtype=torch.uint8, device="cuda")
    d_off = torch.zeros(K + 1, dtype=torch.int64, device="cuda")
    d_len = torch.zeros((K, 256), dtype=torch.uint8, device="cuda")
    codec.encode_raw(d_in.data_ptr(), n, chunk, d_out.data_ptr(), n, d_off.data_ptr(), d_len.data_ptr(), None)
    codec.sync()
    roff = d_off.cpu().numpy().astype(np.uint64)
    assert np.array_equal(off, roff), "chunk offsets"
    assert np.array_equal(lens, d_len.cpu().numpy()), "code lengths"
    assert np.array_equal(payload, d_out[: int(roff[K])].cpu().numpy()), "payload bytes"
    for k in (0, K // 2, K - 1):                          # oracle spot checks
        ref, ln, _ = orc.encode_chunk(data[k * chunk:(k + 1) * chunk])
        assert np.array_equal(payload[int(off[k]):int(off[k + 1])], ref) and np.array_equal(lens[k], ln.astype(np.uint8))
        assert np.array_equal(hist[k].astype(np.uint64), orc.histogram(data[k * chunk:(k + 1) * chunk]))
    sizes = np.diff(off).astype(np.uint32)
    orig = np.array([min(chunk, n - k * chunk) for k in range(K)], dtype=np.uint32)
    back = codec.decode(payload, off[:-1], sizes, orig, lens)       # pipelined decode
    assert np.array_equal(back, data), "round trip"


def test_many_small_chunks_warp_codebook(codec):
    """K >= 1024 chunks take the warp-per-chunk codebook kernel: histogram, code lengths, offsets, payload
    and round trip must still match the oracle chunk by chunk (including one-symbol and ragged chunks)."""
    n, chunk = 1100 * 4096 + 77, 4096
    data = datasets.zipf_stream(n, 3, seed=99).copy()
    data[5 * chunk:6 * chunk] = 0x41                      # a one-symbol chunk
    data[7 * chunk:8 * chunk] = np.arange(chunk, dtype=np.uint32).astype(np.uint8)   # all 256 symbols, equal counts
    check_encode(codec, data, chunk)
    hist = np.zeros((1500, 256), dtype=np.uint32)         # direct histograms: hz_build_codebooks with K >= 1024
    rng = np.random.default_rng(5)
    for k in range(1500):
        m = int(rng.integers(1, 257))
        hist[k, rng.choice(256, m, replace=False)] = rng.integers(1, 1 << int(rng.integers(1, 20)), m)
    _check_codebooks(codec, hist)


@pytest.mark.gpu
@pytest.mark.parametrize("replay", ["warp", "lane0"])
def test_warp_replay_heavy_ties_many_chunks(codec, knob, replay):
    """The heap replay of the codebook kernels (warp_heap_replay: child-preference bits by ballot, path in registers)
    against the oracle's literal PriorityQueue where the tie-breaks decide the lengths: histograms drawn from tiny value
    ranges, sparse alphabets, powers of two, and counts near 2^31 (64-bit heap keys).  1,300 rows take the
    warp-per-chunk kernel (K >= 1024), the first 300 alone the CTA-per-chunk kernel; HZ_CODEBOOK_REPLAY=lane0 runs
    the one-lane literal loops on the same rows."""
    knob("HZ_CODEBOOK_REPLAY", replay)
    rng = np.random.default_rng(20261019)
    hs = []
    for hi in (1, 2, 3, 4, 8, 16, 100):
        for _ in range(80):
            hs.append(rng.integers(0, hi + 1, 256))
            hs.append(rng.integers(1, hi + 1, 256))
    for _ in range(100):
        m = int(rng.integers(2, 257))
        h = np.zeros(256, dtype=np.int64)
        h[rng.choice(256, m, replace=False)] = 1 << rng.integers(0, 6, m)
        hs.append(h)
    for _ in range(60):
        h = rng.integers(0, 3, 256).astype(np.int64)
        h[rng.choice(256, 4, replace=False)] = (1 << 31) // 4 - 200 + rng.integers(0, 3, 4)   # sum stays below 2^31
        hs.append(h)
    while len(hs) < 1300:
        hs.append(rng.permutation(hs[int(rng.integers(0, len(hs)))]))
    hs = np.array(hs, dtype=np.uint32)
    assert (hs.astype(np.int64).sum(axis=1) < (1 << 31)).all()
    _check_codebooks(codec, hs)
    _check_codebooks(codec, hs[:300])


def _chunk_with_counts(rng, counts, size):
    """`size` bytes whose histogram is `counts` (256 ints, sum <= size; symbol 0 absorbs the remainder), shuffled."""
    counts = np.asarray(counts, dtype=np.int64).copy()
    assert counts.sum() <= size
    counts[0] += size - counts.sum()
    b = np.repeat(np.arange(256, dtype=np.uint8), counts)
    rng.shuffle(b)
    return b


def test_lane_codebook_ties_and_long_codes(codec, knob):
    """Streams of thousands of chunks take the lane-per-chunk heap replay (codebook_lane_kernel, 32-bit keys):
    tie-heavy, sparse, one-symbol and long-code chunk histograms, mixed inside the same warps, must give the
    oracle's code lengths, offsets and payload (HZ_CODEBOOK=lane forces the kernel from K >= 1024)."""
    knob("HZ_CODEBOOK", "lane")
    rng = np.random.default_rng(11)
    chunk, K = 2048, 1300
    parts = []
    for k in range(K):
        kind = k % 7
        c = np.zeros(256, dtype=np.int64)
        if kind == 0:
            c[:] = rng.choice([1, 2, 3], 256)
        elif kind == 1:
            m = int(rng.integers(1, 257)); c[rng.choice(256, m, replace=False)] = 1
        elif kind == 2:
            c[rng.choice(256, 40, replace=False)] = 2 ** rng.integers(0, 6, 40)
        elif kind == 3:
            c[int(rng.integers(0, 256))] = chunk                                  # one symbol (symbol 0 adds nothing)
        elif kind == 4:
            c[:] = 8                                                              # all equal: every length 8
        elif kind == 5:
            c[:14] = datasets.fib_like_hist(14)[:14]                              # long codes (max length ~13)
        else:
            c[:] = rng.integers(0, 8, 256)
        parts.append(_chunk_with_counts(rng, c, chunk))
    parts.append(_chunk_with_counts(rng, np.ones(256), 300))                      # ragged last chunk
    check_encode(codec, np.concatenate(parts), chunk)
    # longer codes need bigger chunks: Fibonacci counts up to 24 symbols (max length 23..24), K >= 1024
    chunk, K = 160 * 1024, 1030
    parts = []
    for k in range(K):
        c = np.zeros(256, dtype=np.int64)
        nsym = 10 + k % 15
        c[rng.choice(np.arange(1, 256), nsym, replace=False)] = datasets.fib_like_hist(nsym)[:nsym]
        parts.append(_chunk_with_counts(rng, c, chunk))
    data = np.concatenate(parts)
    payload, off, lens = check_encode(codec, data, chunk)
    assert lens.max() >= 20


@pytest.mark.parametrize("prebuild", ["0", "1"])
def test_decode_tables_built_in_place_or_prebuilt(codec, knob, prebuild):
    """Single-CTA chunks either get their lookup tables from dec_tables_kernel (default while K tables fit the scratch
    cap) or build them inside the sync / write CTAs (HZ_DEC_PREBUILD=0, the path streams of > 150 k chunks take): both
    must decode the same bytes, for small chunks, multi-CTA chunks and a mix of identity and ordinary chunks."""
    knob("HZ_DEC", "legacy")
    knob("HZ_DEC_PREBUILD", prebuild)
    for n, chunk, H in [(3 * MiB + 17, 64 * 1024, 3), (9 * MiB, 4 * MiB, 5), (2 * MiB + 5, 256 * 1024, 8), (700_001, 4096, 1)]:
        data = datasets.zipf_stream(n, H, seed=H + 40)
        check_encode(codec, data, chunk)


def test_two_contexts_one_process(hz):
    """Kernel attributes (opt-in shared memory) are set per context, not once per process."""
    data = datasets.zipf_stream(700_000, 4, seed=5)
    with hz.Codec(0) as a, hz.Codec(0) as b:
        pa, oa, la = a.encode(data, 100_000)
        pb, ob, lb = b.encode(data, 100_000)
        assert np.array_equal(pa, pb) and np.array_equal(oa, ob) and np.array_equal(la, lb)
        K = len(oa) - 1
        orig = np.array([min(100_000, data.size - k * 100_000) for k in range(K)], dtype=np.uint32)
        assert np.array_equal(b.decode(pa, oa[:-1], np.diff(oa).astype(np.uint32), orig, la), data)
