"""Multi-GPU host logic: one process per GPU, chunk-range sharding, host-side offset table.

The path shards by file chunk: every chunk has its own histogram, codebook,
byte-aligned bitstream and footer record (cpu/CpuCompressionService.java:210-261) and chunks are
concatenated in index order (:160-163).  Rank g of G codes chunks [g*K//G, (g+1)*K//G) on its own
GPU with NO data-path collective.  What crosses ranks:

  * every rank's compressed payload and its per-chunk compressed sizes / code lengths / SHA-256 digests, sent to
    rank 0 ONLY (a gather, not an all-gather), which concatenates the payloads in chunk order, assembles the
    offset table with an exclusive scan on the host and writes the footer, byte-identical to the reference's
    (core/CompressionHeader.java:51-85);
  * in the optional GLOBAL-CODEBOOK mode, one all-reduce (sum) of a 256 x int64 histogram
    (2 KiB; NCCL over NVLink on GPUs, gloo in the CPU tests) so that every rank builds the same
    codebook.  That mode is an extension: a valid .dcz any reference decoder accepts, but not
    bit-identical to the reference compressor, which codes every chunk with its own codebook.

Everything here is host logic (numpy / struct / torch.distributed); the per-rank compute goes
through a `Codec` (the C ABI).  `encode_fn` hooks exist so that the CPU test-suite can drive the
same logic with the oracle as the per-rank coder.
"""
import hashlib
import struct

import numpy as np

MAGIC = 0x44435A46          # core/CompressionHeader.java:15
VERSION = 1                 # :16


def bind_to_gpu_numa(device_index):
    """Pin the calling process to the CPUs NVML reports as local to GPU `device_index`, so that pinned host
    buffers allocated afterwards (hz_host_alloc = cudaHostAlloc, first touch) land on that GPU's NUMA node.
    With one process per GPU this keeps every rank's H2D/D2H traffic on its own socket instead of funnelling
    all ranks through the node the launcher happened to start on.  Returns the CPU set, or None when NVML or
    the affinity call is unavailable (nothing is changed then)."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * w + b for w, m in enumerate(words) for b in range(64) if (m >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:
        return None


def chunk_range(K, world, rank):
    """Contiguous chunk range of `rank`: [rank*K//world, (rank+1)*K//world)."""
    return (rank * K) // world, ((rank + 1) * K) // world


def num_chunks(n, chunk_bytes):
    return (n + chunk_bytes - 1) // chunk_bytes          # cpu/CpuCompressionService.java:64


def byte_range(n, chunk_bytes, world, rank):
    """Byte range of the input that `rank` reads."""
    lo, hi = chunk_range(num_chunks(n, chunk_bytes), world, rank)
    return min(n, lo * chunk_bytes), min(n, hi * chunk_bytes)


def assemble_offsets(comp_sizes):
    """Exclusive scan of the per-chunk compressed sizes -> compressedOffset of every chunk
    (cpu/CpuCompressionService.java:137-151).  Host side, as the north star asks."""
    sizes = np.asarray(comp_sizes, dtype=np.uint64)
    off = np.zeros(sizes.size + 1, dtype=np.uint64)
    np.cumsum(sizes, out=off[1:])
    return off


def write_footer(name, orig_size, mtime_ms, chunk_bytes, comp_sizes, orig_sizes, digests, lens, payload_bytes):
    """Footer + 8-byte footer pointer of a .dcz whose payload section is `payload_bytes` long
    (core/CompressionHeader.java:51-85, cpu/CpuCompressionService.java:155-181).  All big-endian."""
    K = len(comp_sizes)
    off = assemble_offsets(comp_sizes)
    nm = name.encode("utf-8")
    g = hashlib.sha256()
    for k in range(K):
        g.update(bytes(digests[k]))                      # SHA-256 over the chunk digests, in index order (:106-109,126)
    out = [struct.pack(">iii", MAGIC, VERSION, len(nm)), nm,
           struct.pack(">qqi", orig_size, mtime_ms, chunk_bytes), g.digest(), struct.pack(">i", K)]
    oo = 0
    for k in range(K):
        out.append(struct.pack(">iqiqi", k, oo, int(orig_sizes[k]), int(off[k]), int(comp_sizes[k])))
        out.append(bytes(digests[k]))
        out.append(np.asarray(lens[k], dtype=">i2").tobytes())
        oo += int(orig_sizes[k])
    out.append(struct.pack(">q", payload_bytes))
    return b"".join(out)


def all_reduce_histogram(hist, group=None):
    """Sum a 256-bin histogram over all ranks (the ONE collective of the path, global-codebook mode).
    `hist` is a torch tensor (int64[256]) on the device the process group works on."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(hist, op=dist.ReduceOp.SUM, group=group)
    return hist


def gather_objects(obj, group=None, dst=0):
    """Per-rank python objects -> list on rank `dst` ONLY (None elsewhere): a rank's compressed payload and its
    per-chunk metadata travel to the rank that writes the file and nowhere else."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return [obj]
    rank = dist.get_rank(group)
    out = [None] * dist.get_world_size(group) if rank == dst else None
    dist.gather_object(obj, out, dst=dist.get_global_rank(group, dst) if group is not None else dst, group=group)
    return out


def sha256_chunks_host(data, chunk_bytes):
    d = memoryview(np.ascontiguousarray(data))
    return [hashlib.sha256(d[o:o + chunk_bytes]).digest() for o in range(0, len(d), chunk_bytes)]


class ShardedCompressor:
    """Compress ONE logical file with G ranks.  Every rank passes ITS byte range of the file
    (byte_range(...)); rank 0 receives the complete .dcz.

    encode_fn(shard, chunk_bytes, fixed_len256 or None) -> (payload uint8[], comp_sizes[Ks], lens[Ks,256])
    hist_fn(shard) -> int64[256] total histogram of the shard      (global-codebook mode only)
    lens_fn(hist int64[256]) -> uint8[256] code lengths             (global-codebook mode only)
    digest_fn(shard, chunk_bytes) -> list of 32-byte SHA-256 digests
    The defaults bind these to a huffb200.Codec (the CUDA path).
    """

    def __init__(self, codec=None, encode_fn=None, hist_fn=None, lens_fn=None, digest_fn=None, device=None):
        self.codec = codec
        self.encode_fn = encode_fn or self._encode_gpu
        self.hist_fn = hist_fn or self._hist_gpu
        self.lens_fn = lens_fn or self._lens_gpu
        self.digest_fn = digest_fn or self._digest_gpu
        self.device = device

    # -- CUDA bindings ------------------------------------------------------------------------
    def _need_codec(self):
        if self.codec is None:
            raise RuntimeError("ShardedCompressor needs a huffb200.Codec: there is no CPU fallback")
        return self.codec

    def _encode_gpu(self, shard, chunk_bytes, fixed_len):
        c = self._need_codec()
        if fixed_len is None:
            payload, off, lens = c.encode(shard, chunk_bytes)
        else:
            payload, off = c.encode_with_lengths(shard, chunk_bytes, fixed_len)
            lens = np.tile(np.asarray(fixed_len, dtype=np.uint8), (len(off) - 1, 1))
        return payload, np.diff(off).astype(np.uint32), lens

    def _hist_gpu(self, shard):
        c = self._need_codec()
        return c.histogram(shard, 1 << 30).astype(np.int64).sum(axis=0) if len(shard) else np.zeros(256, np.int64)

    def _lens_gpu(self, hist):
        c = self._need_codec()
        if int(hist.max()) >= 1 << 32:
            raise ValueError("global histogram bin exceeds 2^32-1 (hz_build_codebooks takes uint32 counts)")
        return c.build_codebooks(hist.astype(np.uint32))[0][0]

    def _digest_gpu(self, shard, chunk_bytes):
        c = self._need_codec()
        return [bytes(d) for d in c.sha256_chunks(shard, chunk_bytes)]

    # -- the sharded pipeline ------------------------------------------------------------------
    def compress(self, shard, total_bytes, chunk_bytes, name, mtime_ms, global_codebook=False, group=None):
        import torch
        import torch.distributed as dist
        rank = dist.get_rank(group) if dist.is_initialized() else 0
        shard = np.ascontiguousarray(np.frombuffer(shard, dtype=np.uint8) if not isinstance(shard, np.ndarray) else shard)
        fixed = None
        if global_codebook:
            h = torch.from_numpy(np.asarray(self.hist_fn(shard), dtype=np.int64).copy())
            if self.device is not None:
                h = h.to(self.device)
            h = all_reduce_histogram(h, group).cpu().numpy()
            fixed = np.asarray(self.lens_fn(h), dtype=np.uint8) if total_bytes else None
        payload, sizes, lens = self.encode_fn(shard, chunk_bytes, fixed)
        Ks = len(sizes)
        orig = [min(chunk_bytes, shard.size - k * chunk_bytes) for k in range(Ks)]
        digests = self.digest_fn(shard, chunk_bytes) if Ks else []
        parts = gather_objects((np.asarray(payload, dtype=np.uint8).tobytes(), list(map(int, sizes)), orig,
                                digests, np.asarray(lens, dtype=np.uint8).reshape(Ks, 256)), group)
        if rank != 0:
            return None
        # rank 0: concatenate payloads in rank (= chunk) order, offset table by exclusive scan on the host
        payload_all = b"".join(p[0] for p in parts)
        sizes_all = [s for p in parts for s in p[1]]
        orig_all = [o for p in parts for o in p[2]]
        dig_all = [d for p in parts for d in p[3]]
        lens_all = np.concatenate([p[4] for p in parts]) if sizes_all else np.zeros((0, 256), np.uint8)
        footer = write_footer(name, total_bytes, mtime_ms, chunk_bytes, sizes_all, orig_all, dig_all, lens_all,
                              len(payload_all))
        return payload_all + footer
