"""Developer tool: SHA-256 of every chunk, GPU kernel (hz_sha256_chunks, one thread per chunk, device-resident
data) against the host's SHA units (one chunk per thread), for the chunk sizes of BASELINE's sweep.
python tools/sha_bench.py [MiB]"""
import os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge
import datasets
hz = ge.load_package()
n = (int(sys.argv[1]) if len(sys.argv) > 1 else 1024) << 20
c = hz.Codec(0)
st = torch.cuda.Stream(); torch.cuda.set_stream(st); c.set_stream(st.cuda_stream)
src = torch.empty(n, dtype=torch.uint8, device="cuda")
c.synth_fill(src.data_ptr(), n, 0, 7, datasets.zipf_qtable(4))
host = src[: 64 << 20].cpu().numpy()
import ctypes as C
L = hz.lib()
out = np.zeros(32, np.uint8)
t0 = time.perf_counter(); L.hz_host_sha256(host.ctypes.data, host.size, out.ctypes.data); t1 = time.perf_counter()
rate1 = host.size / (t1 - t0) / 1e9
cores = len(os.sched_getaffinity(0))
print("# %d MiB device-resident; host SHA-256: %.2f GB/s per thread, %d cores" % (n >> 20, rate1, cores))
print("%9s %8s | %10s | %14s" % ("chunk", "chunks", "GPU GB/s", "host GB/s (est)"))
import hashlib
for ck in (16, 64, 256, 1024, 4096, 16384, 32768):
    chunk = ck << 10
    K = (n + chunk - 1) // chunk
    dig = torch.zeros((K, 32), dtype=torch.uint8, device="cuda")
    f = lambda: c._check(L.hz_sha256_chunks(c._h, src.data_ptr(), n, chunk, dig.data_ptr()))
    f(); c.sync()
    ok = hashlib.sha256(src[:chunk].cpu().numpy()).digest() == bytes(dig[0].cpu().numpy())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(3):
        e0.record(); f(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    print("%8dK %8d | %10.1f | %14.1f  %s" % (ck, K, n / best / 1e6, rate1 * min(cores, K), "ok" if ok else "MISMATCH"))
