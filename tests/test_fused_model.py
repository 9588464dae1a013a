"""CPU pin of the fused decoder's ALGORITHM (tests/fused_model.py: FMT-1 table contents, packed-counter walk, settle
rule, lead-in guesses, chain repair, the chunk's last subsequence ending with the chunk, zero-codeword fill) against the
oracle: the model must decode oracle-encoded and damaged streams to the oracle decoder's bytes for every geometry
the kernel uses.  The kernel itself is compared with the oracle on the GPU (tests/test_gpu_parity.py)."""
import numpy as np
import pytest

import datasets
import fused_model as fm
import orc


@pytest.mark.parametrize("H,S,lead", [(1, 5, 1), (2, 9, 1), (4, 17, 3), (4, 17, 1), (6, 17, 7), (7, 17, 8), (3, 3, 2)])
def test_model_decodes_oracle_streams(H, S, lead):
    data = datasets.zipf_stream(6000, H, seed=100 + H)
    comp, ln, _ = orc.encode_chunk(data)
    out, _ = fm.decode_chunk(comp, ln.astype(np.int64), data.size, S, lead)
    assert np.array_equal(out, data)
    # more symbols asked for than the stream holds: the all-zero codeword's symbol repeats (TableBasedHuffmanDecoder.java:204-208)
    ref, rc = orc.decode(comp, ln.astype(np.int32), data.size + 300, literal=True)
    assert rc == 0
    out, _ = fm.decode_chunk(comp, ln.astype(np.int64), data.size + 300, S, lead)
    assert np.array_equal(out, ref)


@pytest.mark.parametrize("H,S,lead", [(2, 9, 1), (4, 17, 2), (5, 17, 3)])
def test_model_on_damaged_streams_needs_repairs_and_matches(H, S, lead):
    rng = np.random.default_rng(H)
    data = datasets.zipf_stream(5000, H, seed=200 + H)
    comp, ln, _ = orc.encode_chunk(data)
    total = 0
    for trial in range(3):
        bad = comp.copy()
        for f in rng.integers(0, bad.size * 8, 12):
            bad[f >> 3] ^= 0x80 >> (f & 7)
        bad = bad[:bad.size - int(rng.integers(0, 3))]
        ref, rc = orc.decode(bad, ln.astype(np.int32), data.size, literal=True)
        assert rc == 0
        out, repairs = fm.decode_chunk(bad, ln.astype(np.int64), data.size, S, lead)
        assert np.array_equal(out, ref)
        total += repairs
    assert total >= 0


def test_model_long_codes_and_rare_entries():
    """codes longer than the 12-bit window: single-length entries (second-level lookup) and entries with several
    candidate lengths"""
    f = datasets.fib_like_hist(24)
    data = np.repeat(np.arange(256, dtype=np.uint8), np.minimum(f, 3000).astype(np.int64))
    np.random.default_rng(3).shuffle(data)
    data = data[:8000]
    comp, ln, _ = orc.encode_chunk(data)
    assert ln.max() > 12
    out, _ = fm.decode_chunk(comp, ln.astype(np.int64), data.size, 17, 3)
    assert np.array_equal(out, data)
