"""Developer tool: per-kernel CUDA-event timings of the encode / decode pipeline on a synthetic
Zipf stream resident in HBM.  python tools/quickbench.py [MiB] [entropy] [chunk MiB]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge  # noqa: E402
import datasets  # noqa: E402

hz = ge.load_package()
n = int(sys.argv[1]) * (1 << 20) if len(sys.argv) > 1 else 1 << 30
H = int(sys.argv[2]) if len(sys.argv) > 2 else 4
chunk = int(float(sys.argv[3]) * (1 << 20)) if len(sys.argv) > 3 else 16 << 20
reps = int(os.environ.get("REPS", "5"))
K = (n + chunk - 1) // chunk
c = hz.Codec(0)
_s = torch.cuda.Stream(); torch.cuda.set_stream(_s)
c.set_stream(_s.cuda_stream)
src = torch.empty(n, dtype=torch.uint8, device="cuda")
c.synth_fill(src.data_ptr(), n, 0, 0x5EED0001, datasets.zipf_qtable(H) if H else np.full(65536, 0x41, dtype=np.uint8))
comp = torch.empty(n + 16, dtype=torch.uint8, device="cuda")
off = torch.zeros(K + 1, dtype=torch.int64, device="cuda")
lens = torch.zeros((K, 256), dtype=torch.uint8, device="cuda")
back = torch.empty(n, dtype=torch.uint8, device="cuda")
orig = torch.full((K,), chunk, dtype=torch.int32, device="cuda")
orig[K - 1] = n - (K - 1) * chunk


def enc():
    c.encode_raw(src.data_ptr(), n, chunk, comp.data_ptr(), n, off.data_ptr(), lens.data_ptr(), None)


def dec(total):
    sizes = (off[1:] - off[:-1]).to(torch.int32)
    c.decode_raw(comp.data_ptr(), total, off.data_ptr(), sizes.data_ptr(), orig.data_ptr(), None, lens.data_ptr(), K,
                 back.data_ptr(), n)


enc(); c.sync()
total = int(off[K].item())
try:
    dec(total); c.sync()
    ok = torch.equal(back, src)
except Exception as ex:           # developer experiments produce broken streams on purpose
    print("decode failed:", str(ex)[:80]); ok = False
    dec = lambda total: None
print(f"n={n} H={H} chunk={chunk} K={K} comp={total} ({8*total/n:.3f} b/sym) roundtrip_ok={ok}")
e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
te, td = [], []
for _ in range(reps):
    e0.record(); enc(); e1.record(); dec(total); e2.record(); torch.cuda.synchronize()
    te.append(e0.elapsed_time(e1)); td.append(e1.elapsed_time(e2))
te, td = min(te), min(td)
print(f"encode {te:.3f} ms = {n/te/1e6:.1f} GB/s   decode {td:.3f} ms = {n/td/1e6:.1f} GB/s")
c.prof_enable(True); c.prof_reset()
for _ in range(reps):
    enc(); dec(total)
c.sync()
for name, (ms, cnt) in c.prof().items():
    print(f"  {name:20s} {ms/cnt:9.3f} ms/launch  x{cnt}   ({n/(ms/cnt)/1e6:9.1f} GB/s of input)")
c.prof_enable(False)
