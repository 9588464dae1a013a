"""Deterministic inputs shared by the oracle and GPU tests: the reference's own test inputs
 and synthetic streams."""
import os

import numpy as np

import orc

MiB = 1 << 20
GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def reference_cases():
    """(input bytes, file name in footer, chunk bytes, logged .dcz size) — the eleven sizes the
    reference's own test runs logged (app/logs/datacomp-2025-11-12.log, -14.log)."""
    return [
        (b"", "empty.txt", 1 * MiB, 85),
        (b"AAAABBBBCCCCDDDD", "small_input.bin", 16 * MiB, 667),
        (b"A" * 1024, "compressible_input.bin", 16 * MiB, 798),
        (orc.java_random_bytes(42, 1024).tobytes(), "random_input.bin", 16 * MiB, 1672),
        (orc.java_random_bytes(42, 10240).tobytes(), "random.bin", 1 * MiB, 10897),
        (b"Hello World! " * 100, "test.txt", 1 * MiB, 1156),
        (b"Test data for integrity check", "test.txt", 1 * MiB, 671),
        (bytes((ord("A") + (i // 100) % 26) for i in range(512 * 1024)), "speed_test_input.bin", 16 * MiB, 313198),
        (b"A" * (2 * MiB), "test_2mb_input.bin", 16 * MiB, 262810),
        ((np.arange(3 * MiB) % 256).astype(np.uint8).tobytes(), "large.bin", 1 * MiB, 3147529),
        (orc.java_random_bytes(42, 1 * MiB).tobytes(), "test.bin", 16 * MiB, 1049232),
    ]


TEST_INPUT_SHA256 = "fcd8ad1070c6f303848f387385e8f1204be2e2b1a832ea8ce7c12d3ec983c37d"


def fixture_bytes(name):
    """The reference's three binary fixtures: test_small.bin = 2048 x 'A' and test_2mb.bin = 2 MiB x 'A' are
    regenerated (trivially describable); test_input.bin (1 MiB, all 256 byte values, order-0 entropy 7.9998 bits:
    every code length is 8) is the reference's own file, committed as tests/golden/test_input.bin and pinned by
    its SHA-256 (service/gpu/Phase3IntegrationTest.java:148-195 uses it)."""
    if name == "test_small.bin":
        return b"A" * 2048
    if name == "test_2mb.bin":
        return b"A" * (2 * MiB)
    if name == "test_input.bin":
        import hashlib
        with open(os.path.join(GOLDEN_DIR, "test_input.bin"), "rb") as f:
            d = f.read()
        assert hashlib.sha256(d).hexdigest() == TEST_INPUT_SHA256, "tests/golden/test_input.bin is not the reference's fixture"
        return d
    raise KeyError(name)


def zipf_probs(s, nsym=256):
    p = 1.0 / np.arange(1, nsym + 1, dtype=np.float64) ** s
    return p / p.sum()


# Zipf exponents giving order-0 entropy H = 1..8 bits/symbol over 256 symbols
ZIPF_S = {1: 2.9718, 2: 2.1519, 3: 1.7494, 4: 1.4799, 5: 1.2604, 6: 1.0495, 7: 0.7994, 8: 0.0}


def zipf_qtable(entropy_bits, perm_seed=0x5EED):
    """65536-entry quantised inverse CDF for hz_synth_fill: symbol ranks follow a Zipf law with the
    requested entropy; ranks are mapped to byte values by a fixed seeded permutation."""
    p = zipf_probs(ZIPF_S[entropy_bits])
    counts = np.floor(p * 65536).astype(np.int64)
    counts[0] += 65536 - counts.sum()
    perm = np.random.default_rng(perm_seed).permutation(256).astype(np.uint8)
    return np.repeat(perm, counts).astype(np.uint8)


def synth_host(n, seed, qtable, offset=0):
    """Host twin of hz_synth_fill (csrc/hz_synth.cu): qtable[mix64(seed, offset+i) >> 48]."""
    i = np.arange(offset + 1, offset + n + 1, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = np.uint64(seed) + i * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return qtable[(z >> np.uint64(48)).astype(np.int64)]


def zipf_stream(n, entropy_bits, seed=1):
    return synth_host(n, seed, zipf_qtable(entropy_bits))


def fib_like_hist(nsym):
    """Fibonacci-like frequencies -> maximally skewed tree (code lengths up to nsym-1)."""
    f = np.zeros(256, dtype=np.uint64)
    a, b = 1, 1
    for i in range(nsym):
        f[i] = a
        a, b = b, a + b
    return f
