#!/usr/bin/env python
"""Copy the judged evidence of one gpurun profiling call into profiles/ (tracked):
   profiles/<round>_<tag>_bench.json         the bench line of that call
   profiles/<round>_<tag>_launches.txt       ncu launch list: per-kernel count, total / mean device time, share
   profiles/<round>_<tag>_ncu_<kernel>.txt   key metrics + stall breakdown + hottest source lines (ncu --set full)
   profiles/traffic.json                     dram bytes (read+write) per launch of each kernel, read by bench.py
usage: tools/make_profiles.py gpurun_out/<tag> <round>   (needs ncu; no GPU)"""
import collections
import csv
import glob
import io
import json
import os
import subprocess
import sys

src, rnd = sys.argv[1].rstrip("/"), sys.argv[2]
tag = os.path.basename(src)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = os.path.join(ROOT, "profiles")
os.makedirs(out, exist_ok=True)
pre = os.path.join(out, "%s_%s_" % (rnd, tag))
NAMES = {"hist_seg_private": "hist_seg_private", "hist_seg_atomic": "hist_seg_atomic", "hist_seg_lanes": "hist_seg_lanes", "encode_kernel": "encode",
         "dec_sync_kernel": "dec_sync", "dec_write_kernel": "dec_write", "dec_fused_kernel": "dec_fused", "codebook_kernel": "codebook",
         "hist_chain_kernel": "hist_codebook_chain"}

for f in ("bench.json", "bench_ref.json"):
    p = os.path.join(src, f)
    if os.path.exists(p) and os.path.getsize(p):
        open(pre + f, "w").write(open(p).read())

p = os.path.join(src, "launches.csv")
if os.path.exists(p):
    rows = [r for r in csv.reader(l for l in open(p) if not l.startswith("==")) if len(r) > 5]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui], 1e-6)
        name = r[ki].split("(")[0]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(pre + "launches.txt", "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)\n")
        f.write("# command: see tools/gpu_profile.sh; %d launches, %.3f ms total\n" % (sum(a[0] for a in agg.values()), tot))
        f.write("%-34s %8s %12s %12s %8s\n" % ("kernel", "launches", "total_ms", "mean_ms", "share"))
        for name, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("%-34s %8d %12.4f %12.4f %7.1f%%\n" % (name, n, ms, ms / n, 100 * ms / tot))

traffic_path = os.path.join(out, "traffic.json")
traffic = json.load(open(traffic_path)) if os.path.exists(traffic_path) else {}
for rep in sorted(glob.glob(os.path.join(src, "prof_*.ncu-rep"))):
    key = os.path.basename(rep)[len("prof_"):-len(".ncu-rep")]
    txt = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep, "30"],
                         capture_output=True, text=True).stdout
    open(pre + "ncu_%s.txt" % key, "w").write("# ncu --set full --clock-control none --import-source on -k regex:%s (one launch)\n%s" % (key, txt))
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    if len(rows) > 2:
        d = dict(zip(rows[0], rows[2])); u = dict(zip(rows[0], rows[1]))
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
        tb = 0.0
        for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            tb += float(d[m].replace(",", "")) * scale[u[m]]
        kname = d["Kernel Name"].split("(")[0]
        if kname.startswith("void "): kname = kname[5:]              # template instances: "void dec_fused_kernel<24, 1>"
        kname = kname.split("<")[0]
        nbytes = None
        bj = os.path.join(src, "ncu_plain.log")
        traffic[NAMES.get(kname, kname)] = {"dram_bytes_per_launch": tb, "kernel": kname, "capture": "%s_%s" % (rnd, tag),
                                            "ms_under_ncu": float(d["gpu__time_duration.sum"].replace(",", "")) * {"us": 1e-3, "ms": 1, "ns": 1e-6, "s": 1e3}[u["gpu__time_duration.sum"]]}
json.dump(traffic, open(traffic_path, "w"), indent=1, sort_keys=True)
print("wrote", pre + "*", "and", traffic_path)
