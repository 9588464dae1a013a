"""Reduced parity subset for compute-sanitizer (tools/sanitize.sh): every kernel family on small inputs, each result
checked against the CPU oracle.  Exits non-zero on the first mismatch."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge
import datasets, orc
hz = ge.load_package()
c = hz.Codec(0)


def roundtrip(data, chunk, tag):
    data = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data
    payload, off, lens = c.encode(data, chunk)[:3]
    K = len(off) - 1
    for k in range(K):
        ref, ln, _ = orc.encode_chunk(data[k * chunk:(k + 1) * chunk])
        assert np.array_equal(lens[k], ln.astype(np.uint8)) and np.array_equal(payload[int(off[k]):int(off[k + 1])], ref), tag
    orig = np.array([min(chunk, data.size - k * chunk) for k in range(K)], dtype=np.uint32)
    back = c.decode(payload, off[:-1], np.diff(off).astype(np.uint32), orig, lens)
    assert np.array_equal(back, data), tag
    print("ok", tag, flush=True)


for data, name, chunk, expect in datasets.reference_cases()[:8]:
    z = c.compress_buffer(data, chunk, name, 0)
    assert len(z) == expect and z == orc.compress(data, chunk, name, 0) and c.decompress_buffer(z) == bytes(data), name
print("ok reference cases", flush=True)
for mode in ("fused", "legacy"):
    os.environ["HZ_DEC"] = mode; c.reload_knobs()
    roundtrip(datasets.zipf_stream(300_000, 4, seed=1), 100_000, mode + " zipf H4 3 chunks")
    roundtrip(datasets.zipf_stream(1_200_000, 2, seed=2), 1 << 20, mode + " zipf H2 1 MiB chunk (hundreds of units, look-back)")
    roundtrip(datasets.zipf_stream(500_000, 6, seed=3), 500_000, mode + " zipf H6")
    roundtrip(datasets.zipf_stream(2048 * 300, 5, seed=4), 2048, mode + " 300 chunks of 2 KiB")
    rng = np.random.default_rng(9)
    perm = np.concatenate([rng.permutation(256).astype(np.uint8) for _ in range(300)])
    roundtrip(np.concatenate([perm, datasets.zipf_stream(70_000, 3, seed=5)]), 70_000, mode + " identity + ordinary chunk")
    f = datasets.fib_like_hist(30)
    wide = np.repeat(np.arange(256, dtype=np.uint8), np.minimum(f, 40_000).astype(np.int64))
    roundtrip(rng.permutation(wide), 1 << 20, mode + " long codes")
    # damaged stream: bit flips + truncation, the oracle defines the result
    d = datasets.zipf_stream(200_000, 5, seed=6)
    p, off, lens = c.encode(d, 200_000)[:3]
    bad = p.copy()
    for fl in rng.integers(0, bad.size * 8, 20):
        bad[fl >> 3] ^= 0x80 >> (fl & 7)
    bad = bad[:-2]
    ref, rc = orc.decode(bad, lens[0].astype(np.int32), 199_000, literal=False)
    out = c.decode(bad, [0], [bad.size], [199_000], lens[:1])
    assert rc == 0 and np.array_equal(out, ref), mode + " damaged"
    print("ok", mode, "damaged stream", flush=True)
os.environ.pop("HZ_DEC"); c.reload_knobs()
pl, off, l256 = c.encode_global(datasets.zipf_stream(400_000, 3, seed=7), 100_000)
dig = c.sha256_chunks(datasets.zipf_stream(300_000, 6, seed=3), 65536)
print("ok global codebook + sha256", flush=True)
c.close()
print("ALL OK")
