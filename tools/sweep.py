"""Entropy x chunk-size sweep (BASELINE.json configs[4]): encode and decode GB/s of a device-resident
synthetic Zipf stream, CUDA events around the whole encode / decode call.
python tools/sweep.py [MiB] > profiles/<round>_sweep.txt"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge  # noqa: E402
import datasets  # noqa: E402

hz = ge.load_package()
n = int(sys.argv[1]) * (1 << 20) if len(sys.argv) > 1 else 1 << 30
peak = 6444.4
c = hz.Codec(0)
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
c.set_stream(st.cuda_stream)
src = torch.empty(n, dtype=torch.uint8, device="cuda")
comp = torch.empty(n + 16, dtype=torch.uint8, device="cuda")
back = torch.empty(n, dtype=torch.uint8, device="cuda")
print("# %d MiB per point, device-resident, best of 5; roofline = (N + C) / t / %.1f GB/s (measured HBM copy peak)" % (n >> 20, peak))
print("%2s %9s %7s | %9s %7s | %9s %7s | %s" % ("H", "chunk", "b/sym", "enc GB/s", "of peak", "dec GB/s", "of peak", "ok"))
for H in range(1, 9):
    c.synth_fill(src.data_ptr(), n, 0, 0x5EED0001, datasets.zipf_qtable(H))
    for chunk in (64 << 10, 256 << 10, 1 << 20, 4 << 20, 16 << 20):
        K = (n + chunk - 1) // chunk
        off = torch.zeros(K + 1, dtype=torch.int64, device="cuda")
        lens = torch.zeros((K, 256), dtype=torch.uint8, device="cuda")
        orig = torch.full((K,), chunk, dtype=torch.int32, device="cuda")
        orig[K - 1] = n - (K - 1) * chunk
        enc = lambda: c.encode_raw(src.data_ptr(), n, chunk, comp.data_ptr(), n, off.data_ptr(), lens.data_ptr(), None)
        enc(); c.sync()
        C = int(off[K].item())
        sizes = (off[1:] - off[:-1]).to(torch.int32).contiguous()
        dec = lambda: c.decode_raw(comp.data_ptr(), C, off.data_ptr(), sizes.data_ptr(), orig.data_ptr(), None, lens.data_ptr(), K, back.data_ptr(), n)
        dec(); c.sync()
        ok = torch.equal(back, src)
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        te, td = [], []
        for _ in range(5):
            e0.record(); enc(); e1.record(); dec(); e2.record(); torch.cuda.synchronize()
            te.append(e0.elapsed_time(e1)); td.append(e1.elapsed_time(e2))
        te, td = min(te), min(td)
        print("%2d %8dK %7.3f | %9.1f %6.1f%% | %9.1f %6.1f%% | %s" % (H, chunk >> 10, 8 * C / n, n / te / 1e6, 100 * (n + C) / te / 1e6 / peak,
                                                                   n / td / 1e6, 100 * (n + C) / td / 1e6 / peak, ok))
        sys.stdout.flush()
