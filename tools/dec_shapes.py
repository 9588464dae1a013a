"""Developer tool: decode time of a device-resident synthetic Zipf stream for several decoder settings
(environment knobs, re-read per setting) at a list of (entropy, chunk KiB) points.
python tools/dec_shapes.py [MiB] "SETTING;SETTING;..." H:chunkKiB [H:chunkKiB ...]
a SETTING is a comma-separated list of HZ_* assignments, e.g. "HZ_DEC=legacy;HZ_DEC=fused,HZ_FU_WARPS=8;" (empty = defaults)"""
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
settings = sys.argv[2].split(";")
points = [tuple(int(x) for x in a.split(":")) for a in sys.argv[3:]]
reps = int(os.environ.get("REPS", "5"))
c = hz.Codec(0)
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
c.set_stream(st.cuda_stream)
src = torch.empty(n, dtype=torch.uint8, device="cuda")
comp = torch.empty(n + 16, dtype=torch.uint8, device="cuda")
back = torch.empty(n, dtype=torch.uint8, device="cuda")
lastH = None
touched = set()
print("# %d MiB, best of %d, decode GB/s of output bytes (ms per GiB)" % (n >> 20, reps))
print("%-14s" % "point" + "".join(" | %-22s" % (s or "default")[:22] for s in settings))
for H, ck in points:
    if H != lastH:
        c.synth_fill(src.data_ptr(), n, 0, 0x5EED0001, datasets.zipf_qtable(H)); lastH = H
    chunk = ck << 10
    K = (n + chunk - 1) // chunk
    off = torch.zeros(K + 1, dtype=torch.int64, device="cuda")
    lens = torch.zeros((K, 256), dtype=torch.uint8, device="cuda")
    orig = torch.full((K,), chunk, dtype=torch.int32, device="cuda")
    orig[K - 1] = n - (K - 1) * chunk
    c.encode_raw(src.data_ptr(), n, chunk, comp.data_ptr(), n, off.data_ptr(), lens.data_ptr(), None); c.sync()
    C = int(off[K].item())
    sizes = (off[1:] - off[:-1]).to(torch.int32).contiguous()
    dec = lambda: c.decode_raw(comp.data_ptr(), C, off.data_ptr(), sizes.data_ptr(), orig.data_ptr(), None, lens.data_ptr(), K, back.data_ptr(), n)
    row = "H=%d %6dK   " % (H, ck)
    for s in settings:
        for k in touched: os.environ.pop(k, None)
        for kv in filter(None, s.split(",")):
            k, v = kv.split("="); os.environ[k] = v; touched.add(k)
        c.reload_knobs()
        back.zero_()
        dec(); c.sync()
        ok = torch.equal(back, src)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        td = []
        for _ in range(reps):
            e0.record(); dec(); e1.record(); torch.cuda.synchronize()
            td.append(e0.elapsed_time(e1))
        t = min(td)
        row += " | %7.1f (%5.2f) %s     " % (n / t / 1e6, t * (1 << 30) / n, "ok " if ok else "BAD")
    print(row); sys.stdout.flush()
