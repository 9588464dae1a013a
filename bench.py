#!/usr/bin/env python
"""bench.py — Huffman encode/decode throughput of the B200 hot path (BASELINE.json's metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
                    [--size-mib M] [--entropy H] [--chunk-kib C]

One STEP = one pass of the hot path over one synthetic stream: encode the whole stream
(histogram -> codebook -> bit-pack, all chunks) and then decode it again, both through the C ABI
of libhuffb200.so.  `value` = uncompressed GB (1e9 B) coded per second over both directions
(2*N bytes per step; the harmonic mean of the encode and the decode GB/s), stream resident in
HBM.  `e2e` = the same metric with HOST buffers (pinned), host<->device copies in the timed
region.  N>1: one process per GPU (torchrun), every rank codes its own shard of the stream
(weak scaling, no data-path collective; chunks are independent), time = max over ranks.

`--impl reference` times the reference's CPU path (the C++ restatement in oracle/, because no
JVM exists on these boxes) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

MiB = 1 << 20
SEED = 0x5EED0001
METRIC = "huffman_encode_decode_throughput"
UNIT = "GB/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--size-mib", type=int, default=4096, help="stream bytes per GPU (MiB)")
    ap.add_argument("--entropy", type=int, default=4, help="order-0 entropy of the Zipf stream, bits/symbol (1..8)")
    ap.add_argument("--chunk-kib", type=int, default=16 * 1024, help="chunk size in KiB (reference default 16 MiB)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-sample-mib", type=int, default=1024)
    ap.add_argument("--ref-sample-mib", type=int, default=256)
    ap.add_argument("--codebook", default="chunk", choices=["chunk", "global"],
                    help="chunk = one codebook per chunk (reference parity, no collective); global = ONE codebook for "
                         "all ranks: per-rank histograms are all-reduced over NCCL every step (extension mode)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --size-mib per GPU (default); strong: --total-gib split over the GPUs (BASELINE config 4)")
    ap.add_argument("--total-gib", type=int, default=16, help="logical stream of the strong-scaling run (GiB)")
    ap.add_argument("--no-strong", action="store_true", help="skip the extra strong-scaling measurement")
    ap.add_argument("--no-sharded", action="store_true", help="skip the sharded single-file parity check (N > 1)")
    ap.add_argument("--sharded-mib", type=int, default=1024)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    return ap.parse_args()


def workload_name(a):
    return "zipf_H%d_%dMiB_per_gpu_chunk%dKiB" % (a.entropy, a.size_mib, a.chunk_kib)


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy read+write)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def pcie_ceiling(world):
    """Host<->device copy ceiling with `world` GPUs busy at once, measured on this fleet with tools/pcie_peak.py
    (profiles/r02_pcie_concurrent.json), if committed."""
    try:
        with open(os.path.join(ROOT, "profiles", "r02_pcie_concurrent.json")) as f:
            return json.load(f).get(str(world))
    except Exception:
        return None


def ncu_traffic():
    """Per-launch DRAM bytes of each kernel from the committed ncu --set full capture, if any."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f)
    except Exception:
        return {}


# ------------------------------------------------------------------------------------------------
# clocks sampled DURING the timed region
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    BAD = {"hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40, "hw_power_brake": 0x80}
    NOTE = {"sw_power_cap": 0x4, "sync_boost": 0x10, "display_clock": 0x100, "app_clocks": 0x2}

    def __init__(self, cuda_index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        self._h = None
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            self._nv = pynvml
            try:
                uuid = "GPU-" + str(torch.cuda.get_device_properties(cuda_index).uuid)
                self._h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
            except Exception:
                self._h = pynvml.nvmlDeviceGetHandleByIndex(cuda_index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self._h = None

    def _loop(self):
        nv = self._nv
        while not self._stop.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
                try:
                    r = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self._h))
                except Exception:
                    r = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h))
                for name, bit in list(self.BAD.items()) + list(self.NOTE.items()):
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        if self._h is not None:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()

    def stop(self):
        if self._thr is not None:
            self._stop.set()
            self._thr.join()
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ------------------------------------------------------------------------------------------------
# CPU arms (oracle = checker / baseline only; never on the product path)
# ------------------------------------------------------------------------------------------------
def cpu_roundtrip(sample, chunk, threads=0):
    """Literal restatement of the reference CPU path (bit-at-a-time BitOutputStream encode,
    one-bit-peek table decode, java.util.PriorityQueue codebook, chunk-parallel worker pool sized
    max(2, min(nproc, 8)) as cpu/CpuCompressionService.java:42-44).  Returns (enc_s, dec_s, threads)."""
    import orc
    t0 = time.perf_counter()
    comp, sizes, lens, T = orc.encode_chunks_mt(sample, chunk, literal=True, threads=threads)
    t1 = time.perf_counter()
    out, _ = orc.decode_chunks_mt(comp, sizes, lens, sample.size, chunk, literal=True, threads=threads)
    t2 = time.perf_counter()
    if not np.array_equal(out, sample):
        raise RuntimeError("oracle round trip failed")
    return t1 - t0, t2 - t1, T


def host_stream(a, nbytes, offset=0):
    import datasets
    q = datasets.zipf_qtable(a.entropy)
    parts = []
    for o in range(0, nbytes, 64 * MiB):
        parts.append(datasets.synth_host(min(64 * MiB, nbytes - o), SEED, q, offset + o))
    return np.concatenate(parts)


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    chunk = a.chunk_kib * 1024
    nbytes = min(a.ref_sample_mib, a.size_mib) * MiB
    sample = host_stream(a, nbytes)
    for _ in range(a.warmup):
        cpu_roundtrip(sample[: min(nbytes, 4 * chunk)], chunk)
    te = td = 0.0
    T = 0
    t0 = time.perf_counter()
    for _ in range(a.steps):
        e, d, T = cpu_roundtrip(sample, chunk)
        te += e
        td += d
    total = time.perf_counter() - t0
    val = 2.0 * nbytes * a.steps / total / 1e9
    sample_desc = "first %d MiB of the same stream per step (warm-up steps: first %d chunks)" % (nbytes // MiB, 4)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": 1e3 * total / a.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": workload_name(a), "sample": sample_desc,
                   "encode_GBps": nbytes * a.steps / te / 1e9, "decode_GBps": nbytes * a.steps / td / 1e9},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": T, "kind": "port", "sample": sample_desc,
                         "note": "no JVM on the box: the reference's CPU path restated literally in C++ (oracle/), "
                                 "worker pool max(2,min(nproc,8)) as the reference; host has %d cores" % os.cpu_count()},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------
# the B200 arm
# ------------------------------------------------------------------------------------------------
def algorithmic_bytes(name, n, C, K, spc):
    """ALGORITHMIC bytes one launch of kernel `name` must move."""
    if name.startswith("hist"):
        return n
    if name == "encode":
        return n + C
    if name == "dec_sync":
        return C
    if name in ("dec_write", "decode", "dec_fused"):
        return C + n
    if name == "codebook":
        return K * spc * 1024 + K * (256 + 1024)
    return None


class DevicePass:
    """One device-resident workload: `n` bytes of the synthetic stream per rank starting at stream offset `offset`,
    encode + decode through the C ABI on the codec's stream."""

    def __init__(self, env, n, offset, chunk, glob):
        torch, codec, hz = env["torch"], env["codec"], env["hz"]
        import datasets
        self.env, self.n, self.chunk, self.glob = env, n, chunk, glob
        self.K = K = (n + chunk - 1) // chunk
        self.spc = (chunk + hz.SEG_BYTES - 1) // hz.SEG_BYTES
        self.src = torch.empty(n, dtype=torch.uint8, device="cuda")
        codec.synth_fill(self.src.data_ptr(), n, offset, SEED, datasets.zipf_qtable(env["entropy"]))
        cap = n + 16 if not glob else 2 * n + 16            # a shared codebook can expand a shard (bounded by 4x; 2x is ample here)
        self.cap = cap - 16
        self.comp = torch.empty(cap, dtype=torch.uint8, device="cuda")
        self.off = torch.zeros(K + 1, dtype=torch.int64, device="cuda")
        self.lens = torch.zeros((K, 256), dtype=torch.uint8, device="cuda")
        self.back = torch.empty(n, dtype=torch.uint8, device="cuda")
        self.orig = torch.full((K,), chunk, dtype=torch.int32, device="cuda")
        self.orig[K - 1] = n - (K - 1) * chunk
        self.glen = torch.zeros(256, dtype=torch.uint8, device="cuda") if glob else None
        self.enc()
        codec.sync()
        self.C = int(self.off[K].item())
        self.sizes = (self.off[1:] - self.off[:-1]).to(torch.int32).contiguous()
        if glob:
            self.lens.copy_(self.glen.unsqueeze(0).expand(K, 256))
        self.dec()
        codec.sync()
        if not torch.equal(self.back, self.src):
            raise SystemExit("bench.py: decode(encode(x)) != x")

    def enc(self):
        c = self.env["codec"]
        if not self.glob:
            c.encode_raw(self.src.data_ptr(), self.n, self.chunk, self.comp.data_ptr(), self.cap, self.off.data_ptr(),
                         self.lens.data_ptr(), None)
        else:
            # global-codebook mode, all inside the library (hz_encode_global): histogram, device-side reduction,
            # ncclAllReduce of 256 x u64 on the codec's stream, codebook, encode; no torch op on the path
            c.encode_global_raw(self.src.data_ptr(), self.n, self.chunk, self.comp.data_ptr(), self.cap, self.off.data_ptr(),
                                self.glen.data_ptr())

    def dec(self):
        self.env["codec"].decode_raw(self.comp.data_ptr(), self.C, self.off.data_ptr(), self.sizes.data_ptr(),
                                     self.orig.data_ptr(), None, self.lens.data_ptr(), self.K, self.back.data_ptr(), self.n)

    def timed(self, steps, warm, profile, sample_clocks):
        """-> dict(total_ms, enc_ms, dec_ms (max over ranks), enc_all, dec_all (rank 0), prof, launches, clocks)"""
        torch, codec, barrier = self.env["torch"], self.env["codec"], self.env["barrier"]
        for _ in range(max(warm, 3)):
            self.enc()
            self.dec()
        barrier()
        launches0 = codec.launch_count()
        # The timed region runs WITHOUT the library's per-kernel event pairs: an event between the chained histogram /
        # codebook kernel and the encoder launched with programmatic stream serialization makes the encoder wait for the
        # whole kernel before it (same results, no overlap).  The per-kernel table comes from `steps` further, identical
        # steps with the event pairs on (below), so its kernels are timed one after the other.
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * steps + 1)]
        sampler = ClockSampler(self.env["local"]) if sample_clocks else None
        if sampler:
            sampler.start()
        barrier()
        ev[0].record()
        for i in range(steps):
            self.enc()
            ev[2 * i + 1].record()
            self.dec()
            ev[2 * i + 2].record()
        barrier()
        clocks = sampler.stop() if sampler else None
        launches = codec.launch_count() - launches0
        total_ms = ev[0].elapsed_time(ev[-1])
        enc_all = sorted(ev[2 * i].elapsed_time(ev[2 * i + 1]) for i in range(steps))
        dec_all = sorted(ev[2 * i + 1].elapsed_time(ev[2 * i + 2]) for i in range(steps))
        prof = {}
        if profile:
            codec.sync()
            codec.prof_enable(True)
            codec.prof_reset()
            for i in range(steps):
                self.enc()
                self.dec()
            codec.sync()
            prof = codec.prof()
            codec.prof_enable(False)
        codec.sync()
        t = torch.tensor([total_ms, sum(enc_all) / steps, sum(dec_all) / steps], dtype=torch.float64, device="cuda")
        if self.env["world"] > 1:
            self.env["dist"].all_reduce(t, op=self.env["dist"].ReduceOp.MAX)
        total_ms, enc_ms, dec_ms = (float(x) for x in t.cpu())
        return {"total_ms": total_ms, "enc_ms": enc_ms, "dec_ms": dec_ms, "enc_all": enc_all, "dec_all": dec_all,
                "prof": prof, "launches": int(launches), "clocks": clocks}

    def free(self):
        for k in ("src", "comp", "back", "off", "lens", "orig", "sizes", "glen"):
            setattr(self, k, None)
        self.env["torch"].cuda.empty_cache()


def sharded_parity(env, a):
    """ONE logical file compressed by all ranks (chunk-range sharding, per-chunk codebooks = reference parity mode, no
    data-path collective, offset table and footer assembled by rank 0 on the host: cpu/CpuCompressionService.java:
    137-181) must be byte-identical to the single-GPU container and to the oracle's.  -> dict for the JSON line."""
    import hashlib
    import importlib
    torch, codec, hz, dist = env["torch"], env["codec"], env["hz"], env["dist"]
    rank, world = env["rank"], env["world"]
    par = importlib.import_module(hz.__name__ + ".parallel")
    import datasets
    total = min(a.sharded_mib, a.size_mib) * MiB
    chunk = a.chunk_kib * 1024
    lo, hi = par.byte_range(total, chunk, world, rank)
    q = datasets.zipf_qtable(a.entropy)
    d = torch.empty(max(hi - lo, 1), dtype=torch.uint8, device="cuda")
    if hi > lo:
        codec.synth_fill(d.data_ptr(), hi - lo, lo, SEED ^ 0x5A, q)
    shard = d[: hi - lo].cpu().numpy()
    t0 = time.perf_counter()
    blob = par.ShardedCompressor(codec, device=torch.device("cuda", env["local"])).compress(
        shard, total, chunk, "bench.bin", 1_700_000_000_000)
    t_sharded = time.perf_counter() - t0
    if rank != 0:
        return None
    import orc
    full = torch.empty(total, dtype=torch.uint8, device="cuda")
    codec.synth_fill(full.data_ptr(), total, 0, SEED ^ 0x5A, q)
    host = full.cpu().numpy()
    del full
    one = codec.compress_buffer(host, chunk, "bench.bin", 1_700_000_000_000)
    ref = orc.compress(host, chunk, "bench.bin", 1_700_000_000_000, literal=False)
    h = [hashlib.sha256(x).hexdigest() for x in (blob, one, ref)]
    ok = h[0] == h[1] == h[2]
    if not ok:
        raise SystemExit("bench.py: sharded .dcz differs (sharded %s, one GPU %s, oracle %s)" % tuple(x[:16] for x in h))
    return {"sharded_parity": True, "logical_bytes": total, "ranks": world, "dcz_bytes": len(blob), "dcz_sha256": h[0],
            "compared_with": "single-GPU hz_compress_buffer and the CPU oracle (orc.compress)", "wall_s": t_sharded}


def run_b200(a):
    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 arm has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    hz = ge.load_package()
    import importlib
    numa_cpus = importlib.import_module(hz.__name__ + ".parallel").bind_to_gpu_numa(local)   # pinned buffers on the GPU's socket
    codec = hz.Codec(local)
    stream = torch.cuda.Stream()          # a real (non-default) stream: the codec, torch and the events share it
    torch.cuda.set_stream(stream)
    codec.set_stream(stream.cuda_stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    env = {"torch": torch, "dist": dist, "codec": codec, "hz": hz, "rank": rank, "world": world, "local": local,
           "entropy": a.entropy, "barrier": barrier}
    glob = a.codebook == "global"
    if glob and world > 1:
        # the library owns the communicator of its one collective; the 128-byte id travels over torch.distributed ONCE
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(hz.Codec.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        codec.comm_init(bytes(idt.cpu().numpy().tobytes()), world, rank)

    chunk = a.chunk_kib * 1024
    strong_main = a.scaling == "strong"
    n = a.size_mib * MiB if not strong_main else (a.total_gib << 30) // world
    wp = DevicePass(env, n, rank * n, chunk, glob)
    K, spc, C = wp.K, wp.spc, wp.C
    src = wp.src
    if rank == 0 and not glob:
        # bit-exactness spot check of chunk 0 against the CPU oracle (checker use only)
        import orc
        c0 = src[:min(n, chunk)].cpu().numpy()
        ref, ln, _ = orc.encode_chunk(c0)
        got = wp.comp[: int(wp.off[1].item())].cpu().numpy()
        if not (np.array_equal(got, ref) and np.array_equal(wp.lens[0].cpu().numpy(), ln.astype(np.uint8))):
            raise SystemExit("bench.py: chunk 0 differs from the oracle")

    m = wp.timed(a.steps, a.warmup, True, rank == 0)
    total_ms, enc_ms, dec_ms, enc_all, dec_all, prof, launches, clocks = (m[k] for k in (
        "total_ms", "enc_ms", "dec_ms", "enc_all", "dec_all", "prof", "launches", "clocks"))
    ms_per_step = total_ms / a.steps
    value = 2.0 * n * world / (ms_per_step * 1e6)

    # ---- end to end through the C ABI with host buffers (pinned), copies inside the timed region
    e2e = None
    if not a.no_e2e and not glob:
        orig = wp.orig
        h_src = torch.empty(n, dtype=torch.uint8).pin_memory()
        h_src.copy_(src)
        h_comp = torch.empty(n + 16, dtype=torch.uint8).pin_memory()
        h_back = torch.empty(n, dtype=torch.uint8).pin_memory()
        h_off = np.zeros(K + 1, dtype=np.uint64)
        h_lens = np.zeros((K, 256), dtype=np.uint8)
        h_orig = orig.cpu().numpy().astype(np.uint32)

        def e2e_step():
            codec.encode_raw(h_src.data_ptr(), n, chunk, h_comp.data_ptr(), n, h_off, h_lens, None)
            csz = np.diff(h_off).astype(np.uint32)
            codec.decode_raw(h_comp.data_ptr(), int(h_off[K]), h_off, csz, h_orig, None, h_lens, K, h_back.data_ptr(), n)
            codec.sync()

        e2e_step()
        if not torch.equal(h_back, h_src):
            raise SystemExit("bench.py: e2e round trip failed")
        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(a.e2e_steps):
            e2e_step()
        barrier()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e_s = float(dt.item()) / a.e2e_steps
        meta = (K + 1) * 8 + K * 256
        e2e = {"value": 2.0 * n * world / e2e_s / 1e9, "unit": UNIT, "ms_per_step": 1e3 * e2e_s, "steps": a.e2e_steps,
               "h2d_bytes_per_step": n + C + 2 * meta + 8 * K, "d2h_bytes_per_step": C + n + meta,
               "api": "hz_encode + hz_decode with pinned host buffers",
               "host_numa_binding": ("rank pinned to the %d CPUs NVML reports local to its GPU before the pinned buffers are allocated"
                                     % len(numa_cpus)) if numa_cpus else "none (NVML affinity unavailable)"}
        pc = pcie_ceiling(world)
        if pc:
            # the two calls of a step do not overlap each other (decode consumes encode's output): the encode call is bound
            # by its busier direction, max(N up, C down), and so is the decode call, max(C up, N down); with both directions
            # busy the host sustains `duplex_GBps_per_direction` per GPU when `world` GPUs copy at once (tools/pcie_peak.py)
            bound_s = 2.0 * max(n, C) / (pc["duplex_GBps_per_direction"] * 1e9)
            e2e["host_link_ceiling"] = dict(pc, bound_ms_per_step=1e3 * bound_s, frac_of_ceiling=bound_s / e2e_s,
                                            bound_is="2 * max(N, C) / duplex rate: the busier direction of each of the two calls")
        del h_src, h_comp, h_back

    cpu_sample = None
    if rank == 0 and not a.no_cpu and world == 1:
        cpu_sample = src[:min(a.cpu_sample_mib, a.size_mib) * MiB].cpu().numpy()
    wp.free()
    del src

    # ---- BASELINE config 4: ONE logical stream of --total-gib split over the ranks (strong scaling), same kernels
    strong = None
    if not a.no_strong and not strong_main:
        ns = (a.total_gib << 30) // world
        sp = DevicePass(env, ns, rank * ns, chunk, glob)
        sm = sp.timed(max(3, a.steps // 2), 3, False, False)
        st_steps = max(3, a.steps // 2)
        strong = {"scaling": "strong", "total_bytes": ns * world, "bytes_per_gpu": ns, "steps": st_steps,
                  "ms_per_step": sm["total_ms"] / st_steps, "value": 2.0 * ns * world / (sm["total_ms"] / st_steps * 1e6), "unit": UNIT,
                  "encode_GBps": ns * world / (sm["enc_ms"] * 1e6), "decode_GBps": ns * world / (sm["dec_ms"] * 1e6),
                  "note": "%d GiB logical stream, %d GiB per GPU; efficiency(N) = value(N) / (N * value(1))" % (a.total_gib, ns >> 30)}
        sp.free()

    # ---- N > 1: all ranks compress ONE file; rank 0's container must equal the 1-GPU and the oracle's bytes
    sharded = None
    if world > 1 and not a.no_sharded and not glob:
        sharded = sharded_parity(env, a)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline: both directions against the measured HBM peak + the dominant kernel + per-kernel table
    peak, peak_src = measured_peak()
    traffic = ncu_traffic()
    kernels = []
    for name, (ms, cnt) in prof.items():
        per = ms / max(cnt, 1)
        ab = algorithmic_bytes(name, n, C, K, spc)
        lps = cnt / a.steps                                   # (the two-part encode launches histogram / codebook / encoder twice)
        kernels.append({"name": name, "launches_per_step": lps, "ms_per_launch": per,
                        "share_of_step": ms / a.steps / ms_per_step,
                        "algorithmic_bytes": ab,                # of all launches of one step together
                        "achieved_GBps": (ab / (per * lps * 1e6)) if ab and per > 0 else None})
    kernels.sort(key=lambda k: -k["share_of_step"])
    dom = next((k for k in kernels if k["algorithmic_bytes"]), None)
    stages = {
        "encode": {"GBps": n / (enc_ms * 1e6), "ms": enc_ms, "algorithmic_bytes": n + C,
                   "achieved": (n + C) / (enc_ms * 1e6), "frac": (n + C) / (enc_ms * 1e6) / peak,
                   "frac_of_nominal_8000": (n + C) / (enc_ms * 1e6) / 8000.0,
                   "ms_min_rank0": enc_all[0], "ms_median_rank0": enc_all[len(enc_all) // 2]},
        "decode": {"GBps": n / (dec_ms * 1e6), "ms": dec_ms, "algorithmic_bytes": C + n,
                   "achieved": (C + n) / (dec_ms * 1e6), "frac": (C + n) / (dec_ms * 1e6) / peak,
                   "frac_of_nominal_8000": (C + n) / (dec_ms * 1e6) / 8000.0,
                   "ms_min_rank0": dec_all[0], "ms_median_rank0": dec_all[len(dec_all) // 2]},
    }
    low = min(stages, key=lambda d: stages[d]["frac"])
    roofline = {"bound": "hbm", "achieved": stages[low]["achieved"], "peak": peak, "unit": "GB/s", "frac": stages[low]["frac"],
                "frac_is": "the LOWER of the two directions (%s): algorithmic bytes (N + C) / stage time / peak; the north star "
                           "states its targets per direction (encode >= 0.50, decode >= 0.40)" % low,
                "stages": {d: {"frac": stages[d]["frac"], "ms": stages[d]["ms"], "achieved": stages[d]["achieved"],
                               "GBps_of_uncompressed_bytes": stages[d]["GBps"]} for d in stages},
                "peak_source": peak_src, "traffic": None}
    if dom:
        roofline["kernel"] = {"name": dom["name"], "achieved": dom["achieved_GBps"], "frac": dom["achieved_GBps"] / peak,
                              "algorithmic_bytes_per_launch": dom["algorithmic_bytes"] / dom["launches_per_step"], "ms_per_launch": dom["ms_per_launch"],
                              "frac_of_nominal_8000": dom["achieved_GBps"] / 8000.0}
        if a.size_mib == 4096 and not strong_main:
            roofline["traffic"] = (traffic.get(dom["name"]) or {}).get("dram_bytes_per_launch")
            roofline["traffic_source"] = "profiles/traffic.json (ncu --set full, default 4 GiB workload), dominant kernel per launch"

    cpu = None
    if cpu_sample is not None:
        nb = cpu_sample.size
        e, d, T = cpu_roundtrip(cpu_sample, chunk)
        cpu = {"value": 2.0 * nb / (e + d) / 1e9, "unit": UNIT, "cores": T, "kind": "port",
               "sample": "first %d MiB of the benchmark stream, one encode+decode pass" % (nb // MiB),
               "encode_GBps": nb / e / 1e9, "decode_GBps": nb / d / 1e9, "host_cores": os.cpu_count()}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": a.scaling, "vs_baseline": None, "dtype": "u8",
        "data": "synthetic",
        "config": {"workload": workload_name(a) if not strong_main else "zipf_H%d_%dGiB_total_chunk%dKiB" % (a.entropy, a.total_gib, a.chunk_kib),
                   "bytes_per_gpu": n, "chunk_bytes": chunk, "chunks_per_gpu": K,
                   "compressed_bytes_per_gpu": C, "bits_per_symbol": 8.0 * C / n, "seed": SEED,
                   "value_definition": "2*N*n_gpus / step time; step = encode(all chunks) then decode(all chunks), device-resident",
                   "l2": "inputs (%d MiB) are larger than the 126 MB L2; no flush between iterations" % (n // MiB),
                   "codebooks": "per chunk (reference parity mode)" if not glob else
                                "ONE global codebook (hz_encode_global): per-rank histograms all-reduced (sum of 256 x u64) "
                                "by ncclAllReduce on the codec's stream every step"},
        "stages": stages, "roofline": roofline, "kernels": kernels,
        "kernels_from": "%d further steps of the same loop with the library's CUDA-event pair around every launch (kernels then run "
                        "one after the other: shares can sum to > 1; the timed steps run without the pairs because an event between "
                        "the chained histogram/codebook kernel and its programmatically dependent encoder removes their overlap)" % a.steps,
        "cpu_baseline": cpu, "e2e": e2e,
        "strong_scaling": strong, "sharded": sharded,
        "gpu_launches": int(launches), "clocks": clocks,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    a = parse_args()
    # stdout carries exactly ONE JSON line: everything else written to fd 1 (NCCL prints its version banner
    # there) is sent to stderr; the JSON goes to the saved descriptor
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = real
    if a.impl == "reference":
        return run_reference(a)
    return run_b200(a)


if __name__ == "__main__":
    sys.exit(main())
