"""World-size-2 (and 3) CPU tests of the multi-GPU host logic (huffb200.parallel) over gloo:
chunk-range sharding, metadata gather, host-side offset table / footer assembly, and the one
collective of the path (histogram all-reduce, global-codebook mode).  The per-rank coder is the
CPU oracle here (test infrastructure); on GPUs the same class binds to the CUDA codec
(tests/test_gpu_parity.py::test_sharded_*)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _oracle_fns():
    import orc

    def encode_fn(shard, chunk, fixed):
        K = (shard.size + chunk - 1) // chunk
        parts, sizes, lens = [], [], np.zeros((K, 256), np.uint8)
        for k in range(K):
            c = shard[k * chunk:(k + 1) * chunk]
            if fixed is None:
                p, ln, _ = orc.encode_chunk(c)
            else:
                ln = np.asarray(fixed, dtype=np.int32)
                cd, _ = orc.canonical_codes(ln)
                p = orc.encode(c, ln, cd)
            parts.append(p); sizes.append(p.size); lens[k] = ln
        return (np.concatenate(parts) if parts else np.zeros(0, np.uint8)), np.asarray(sizes, np.uint32), lens

    def hist_fn(shard):
        return orc.histogram(shard).astype(np.int64)

    def lens_fn(hist):
        ln, mx = orc.code_lengths(hist.astype(np.uint64))
        assert mx >= 0
        return ln.astype(np.uint8)

    return encode_fn, hist_fn, lens_fn


def _worker(rank, world, port, n, chunk, q):
    import torch.distributed as dist
    import __graft_entry__ as ge
    import datasets
    import orc
    hz = ge.load_package()
    import importlib
    par = importlib.import_module("huffb200.parallel")
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        data = datasets.zipf_stream(n, 3, seed=11) if n else np.zeros(0, np.uint8)
        lo, hi = par.byte_range(n, chunk, world, rank)
        enc, hist, lens = _oracle_fns()
        sc = par.ShardedCompressor(encode_fn=enc, hist_fn=hist, lens_fn=lens, digest_fn=par.sha256_chunks_host)
        out = sc.compress(data[lo:hi], n, chunk, "shard.bin", 1234567)
        glob = sc.compress(data[lo:hi], n, chunk, "shard.bin", 1234567, global_codebook=True)
        if rank == 0:
            ref = orc.compress(data, chunk, "shard.bin", 1234567)
            ok_parity = out == ref
            back = orc.decompress(glob)
            ok_global = back == data.tobytes()
            # global mode: every chunk carries the same code lengths (one codebook for the file)
            q.put((ok_parity, ok_global, len(out), len(ref)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n,chunk", [(2, 1_000_003, 100_000), (2, 300_000, 1 << 20), (3, 777_777, 65_536), (2, 0, 4096)])
def test_sharded_compress_matches_single_process_oracle(world, n, chunk):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, chunk, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    ok_parity, ok_global, got, want = q.get(timeout=10)
    assert got == want
    assert ok_parity, "sharded .dcz differs from the single-process reference container"
    assert ok_global, "global-codebook .dcz does not decode to the input"


def test_chunk_ranges_partition():
    import __graft_entry__ as ge
    ge.load_package()
    import importlib
    par = importlib.import_module("huffb200.parallel")
    for K in (0, 1, 7, 8, 64, 1000):
        for G in (1, 2, 3, 4, 8):
            r = [par.chunk_range(K, G, g) for g in range(G)]
            assert r[0][0] == 0 and r[-1][1] == K
            assert all(r[i][1] == r[i + 1][0] for i in range(G - 1))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1
    off = par.assemble_offsets([5, 0, 7])
    assert off.tolist() == [0, 5, 5, 12]
