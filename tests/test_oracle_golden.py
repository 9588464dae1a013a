"""CPU-only: pins the oracle (oracle/huff_oracle.cpp) against every known answer the reference
tree holds for the hot path.  No GPU, no product code."""
import hashlib
import json
import os

import numpy as np
import pytest

import datasets
import orc


@pytest.mark.parametrize("idx", range(11))
def test_logged_dcz_sizes(idx):
    data, name, chunk, expect = datasets.reference_cases()[idx]
    z = orc.compress(data, chunk, name, 0, literal=True)
    assert len(z) == expect                       # the size the reference's own run logged
    assert orc.compress(data, chunk, name, 0, literal=False) == z     # literal and fast coder agree
    assert orc.decompress(z, literal=True) == data
    assert orc.decompress(z, literal=False) == data


def test_size_identity():
    # out = payload + (68 + len(name)) + 572*K + 8   
    data, name, chunk, expect = datasets.reference_cases()[9]
    K = 3
    z = orc.compress(data, chunk, name, 0)
    payload = int.from_bytes(z[-8:], "big")
    assert len(z) == payload + 68 + len(name) + 572 * K + 8


def test_java_random_first_bytes():
    # new Random(42).nextInt() == -1170105035 -> bytes little-endian-first
    b = orc.java_random_bytes(42, 8)
    assert int.from_bytes(b[:4].tobytes(), "little", signed=True) == -1170105035
    assert int.from_bytes(b[4:8].tobytes(), "little", signed=True) == 234785527


def test_histogram_known_answers():
    # CpuFrequencyServiceTest.java:25-35, :38-49, :70-80, :83-91
    h = orc.histogram(bytes([0, 1, 2, 1, 0, 1]))
    assert (h[0], h[1], h[2]) == (2, 3, 1)
    assert np.all(orc.histogram(bytes(range(256))) == 1)
    d = bytes(5 if i < 50 else 10 for i in range(100))
    h = orc.histogram(d[25:75])
    assert h[5] == 25 and h[10] == 25
    h = orc.histogram(np.array([-1, -2, -3, -1], dtype=np.int8).view(np.uint8))
    assert (h[255], h[254], h[253]) == (2, 1, 1)


def test_codebook_structure():
    # CanonicalHuffmanTest.java:12-27 uniform, :30-45 skewed, :48-57 single, :60-66 empty, :69-94 canonical
    ln, mx = orc.code_lengths(np.full(256, 100))
    assert np.all(ln == 8)
    f = np.ones(256, dtype=np.uint64); f[0], f[1], f[2] = 1000, 500, 250
    ln, _ = orc.code_lengths(f)
    assert ln[0] <= ln[255]
    f = np.zeros(256, dtype=np.uint64); f[42] = 1000
    ln, mx = orc.code_lengths(f)
    assert ln[42] == 1 and mx == 1 and ln.sum() == 1
    ln, mx = orc.code_lengths(np.zeros(256))
    assert mx == 0 and not ln.any()
    ln, _ = orc.code_lengths(np.arange(1, 257))
    code, _ = orc.canonical_codes(ln)
    for l in range(1, 17):
        c = code[ln == l]
        assert np.all(np.diff(c.astype(np.int64)) == 1)


def test_property_prefix_free_and_kraft():
    # HuffmanPropertyTest.java:12-38,41-66,69-78 with its generator (:81-92): freq in [0,1000]^256
    rng = np.random.default_rng(123)
    for _ in range(200):
        f = rng.integers(0, 1001, 256).astype(np.uint64)
        f[rng.integers(0, 256)] += 1
        ln, mx = orc.code_lengths(f)
        code, _ = orc.canonical_codes(ln)
        nz = f > 0
        assert np.all(ln[nz] > 0) and np.all(ln[~nz] == 0)
        assert ln[np.argmax(f)] <= ln[nz][np.argmin(f[nz])]
        if nz.sum() >= 2:
            assert sum(2.0 ** -int(l) for l in ln[nz]) == 1.0
        keys = set((int(l), int(c)) for l, c in zip(ln[nz], code[nz]))
        assert len(keys) == nz.sum()


def test_msb_first_merge_examples():
    # ReductionBasedEncodingTest.java:27-66: "shift a left by len_b, OR in b"; :79-114: 8 -> 1
    ln = np.zeros(256, dtype=np.int32); cd = np.zeros(256, dtype=np.uint32)
    ln[0], cd[0] = 1, 0b0
    ln[1], cd[1] = 2, 0b10
    ln[2], cd[2] = 3, 0b110
    ln[3], cd[3] = 3, 0b111
    out = orc.encode(bytes([1, 2, 3]), ln, cd, literal=True)       # 10 110 111 -> 10110111
    assert out.tobytes() == bytes([0b10110111])
    out = orc.encode(bytes([0, 1, 0]), ln, cd)                    # 0 10 0 + 4 pad bits
    assert out.tobytes() == bytes([0b01000000])


def test_survey_vectors():
    # SURVEY.md §8c "survey-derived vectors" (restatement outputs consistent with the logged sizes)
    p, ln, cd = orc.encode_chunk(b"Hello World! " * 100)
    got = [(s, int(ln[s]), int(cd[s])) for s in range(256) if ln[s]]
    assert got == [(32, 3, 4), (33, 4, 10), (72, 4, 11), (87, 4, 12), (100, 4, 13), (101, 4, 14),
                   (108, 2, 0), (111, 2, 1), (114, 4, 15)]
    assert len(p) == 500 and hashlib.sha256(p).hexdigest().startswith("9e12ee6558ca8ff649b56428")
    p, _, _ = orc.encode_chunk(b"AAAABBBBCCCCDDDD")
    assert p.tobytes() == bytes.fromhex("0055aaff")
    p, ln, _ = orc.encode_chunk(orc.java_random_bytes(42, 1024))
    assert ln.max() == 11 and len(p) == 1008 and hashlib.sha256(p).hexdigest().startswith("625239accf7f8bf75191a988")
    p, ln, _ = orc.encode_chunk(orc.java_random_bytes(42, 10240))
    assert ln.max() == 9 and len(p) == 10239 and hashlib.sha256(p).hexdigest().startswith("75323787606b5a9e82f64c0e")
    p, ln, _ = orc.encode_chunk(bytes((ord("A") + (i // 100) % 26) for i in range(512 * 1024)))
    assert ln.max() == 5 and len(p) == 312530 and hashlib.sha256(p).hexdigest().startswith("9c3e74aea1d6c05576648faf")


def test_fixtures():
    # test_small.bin / test_2mb.bin: single symbol -> 1-bit code 0, payload all zero bytes
    p, ln, _ = orc.encode_chunk(datasets.fixture_bytes("test_small.bin"))
    assert ln[65] == 1 and ln.sum() == 1 and p.tobytes() == bytes(256)
    z = orc.compress(datasets.fixture_bytes("test_small.bin"), 32 * datasets.MiB, "test_small.bin", 0)
    assert len(z) == 256 + (68 + 14) + 572 + 8
    z = orc.compress(datasets.fixture_bytes("test_2mb.bin"), 32 * datasets.MiB, "test_2mb.bin", 0)
    assert len(z) == 262144 + 80 + 572 + 8
    z = orc.compress(datasets.fixture_bytes("test_2mb.bin"), 1 * datasets.MiB, "test_2mb.bin", 0)
    assert int.from_bytes(z[-8:], "big") == 2 * 131072
    # test_input.bin (the reference's file itself): every length 8 -> canonical code[s] == s -> payload == input
    d = np.frombuffer(datasets.fixture_bytes("test_input.bin"), dtype=np.uint8)
    p, ln, _ = orc.encode_chunk(d)
    assert (ln == 8).all() and np.array_equal(p, d)
    z = orc.compress(d, 16 * datasets.MiB, "test_input.bin", 0)
    assert len(z) == 1048576 + 68 + 14 + 572 + 8 and z[:1048576] == d.tobytes()
    assert orc.decompress(z) == d.tobytes()


def test_literal_and_fast_decoders_agree_on_damaged_streams():
    """The GPU parity tests compare damaged streams with the oracle's word-at-a-time decoder; this pins that decoder to
    the literal restatement of TableBasedHuffmanDecoder (one-bit peeks, 10-bit table, bit-by-bit fallback:
    core/TableBasedHuffmanDecoder.java:103-152,180-231) on exactly such inputs: flipped bits, truncated payloads
    (zero bits past the end, :204-208) and a short orig_size, for short and long codes."""
    rng = np.random.default_rng(2026)
    n = 60_000
    for H in (1, 2, 4, 5, 7):
        data = datasets.zipf_stream(n, H, seed=H + 90)
        p, ln, _ = orc.encode_chunk(data)
        ln = ln.astype(np.int32)
        for trial in range(6):
            bad = p.copy()
            for f in rng.integers(0, bad.size * 8, 12):
                bad[f >> 3] ^= 0x80 >> (f & 7)
            bad = bad[: bad.size - int(rng.integers(0, 4))]
            n_out = n - int(rng.integers(0, 3)) * 500
            a, ra = orc.decode(bad, ln, n_out, literal=True)
            b, rb = orc.decode(bad, ln, n_out, literal=False)
            assert ra == rb and np.array_equal(a, b), "H=%d trial %d" % (H, trial)
    # an incomplete code (one symbol): a stray 1 bit is an error for both
    one = np.zeros(256, dtype=np.int32); one[65] = 1
    comp = np.zeros(8, dtype=np.uint8); comp[3] = 0x10
    assert orc.decode(comp, one, 40, literal=True)[1] == orc.decode(comp, one, 40, literal=False)[1] != 0


def test_decode_literal_vs_fast_and_long_codes():
    rng = np.random.default_rng(5)
    # long codes (> 10 bits: the reference's fallback path) from a Fibonacci-like histogram
    f = datasets.fib_like_hist(24)
    ln, mx = orc.code_lengths(f)
    assert mx == 23
    cd, _ = orc.canonical_codes(ln)
    p = f[:24] / f[:24].sum()
    data = rng.choice(24, size=20000, p=p).astype(np.uint8)
    data[:24] = np.arange(24)                                    # every symbol at least once
    enc = orc.encode(data, ln, cd, literal=True)
    assert np.array_equal(enc, orc.encode(data, ln, cd, literal=False))
    a, rc1 = orc.decode(enc, ln, data.size, literal=True)
    b, rc2 = orc.decode(enc, ln, data.size, literal=False)
    assert rc1 == 0 and rc2 == 0 and np.array_equal(a, data) and np.array_equal(b, data)


def test_decode_error_position():
    # single-symbol chunk: code '0'; a 1 bit matches nothing -> "Huffman decode error at position i"
    ln = np.zeros(256, dtype=np.int32); ln[65] = 1
    comp = np.array([0b00010000], dtype=np.uint8)
    for lit in (True, False):
        _, rc = orc.decode(comp, ln, 8, literal=lit)
        assert rc == -(3 + 1)


def test_code_too_long_is_an_error():
    ln, mx = orc.code_lengths(datasets.fib_like_hist(40))       # depth 39 > 32: reference throws
    assert mx == -1


def test_golden_file_matches_oracle():
    """tests/golden/reference_cases.json was produced by tests/golden/make_golden.py from this
    oracle; it freezes payload hashes + code lengths so an oracle regression is caught too."""
    path = os.path.join(datasets.GOLDEN_DIR, "reference_cases.json")
    golden = json.load(open(path))
    cases = datasets.reference_cases()
    assert len(golden) == len(cases)
    for g, (data, name, chunk, expect) in zip(golden, cases):
        z = orc.compress(data, chunk, name, 0)
        assert g["dcz_size"] == expect == len(z)
        assert g["dcz_sha256"] == hashlib.sha256(z).hexdigest()
