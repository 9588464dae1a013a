"""Developer tool: per-kernel CUDA-event times (library profiler, hz_prof_*) of encode + decode for a list of
(entropy, chunk KiB) points on a device-resident synthetic Zipf stream.
python tools/kprof.py [MiB] H:chunkKiB [H:chunkKiB ...]      e.g.  tools/kprof.py 1024 4:64 4:256 8:16384"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge  # noqa: E402
import datasets  # noqa: E402

hz = ge.load_package()
n = int(sys.argv[1]) * (1 << 20)
points = [tuple(int(x) for x in a.split(":")) for a in sys.argv[2:]]
reps = int(os.environ.get("REPS", "5"))
c = hz.Codec(0)
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
c.set_stream(st.cuda_stream)
src = torch.empty(n, dtype=torch.uint8, device="cuda")
comp = torch.empty(n + 16, dtype=torch.uint8, device="cuda")
back = torch.empty(n, dtype=torch.uint8, device="cuda")
lastH = None
for H, ck in points:
    if H != lastH:
        c.synth_fill(src.data_ptr(), n, 0, 0x5EED0001, datasets.zipf_qtable(H)); lastH = H
    chunk = ck << 10
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
    for _ in range(reps):
        e0.record(); enc(); e1.record(); dec(); e2.record(); torch.cuda.synchronize()
        te.append(e0.elapsed_time(e1)); td.append(e1.elapsed_time(e2))
    c.prof_enable(True); c.prof_reset()
    for _ in range(reps):
        enc(); dec()
    c.sync()
    pr = c.prof(); c.prof_enable(False)
    print("H=%d chunk=%dK K=%d b/sym=%.3f ok=%s  encode %.3f ms (%.0f GB/s)  decode %.3f ms (%.0f GB/s)" % (
        H, ck, K, 8 * C / n, ok, min(te), n / min(te) / 1e6, min(td), n / min(td) / 1e6))
    print("   " + "  ".join("%s %.3f" % (k, v[0] / max(1, v[1])) for k, v in pr.items()))
    sys.stdout.flush()
