#!/usr/bin/env python
"""Per CUDA source line: executed warp instructions and stall samples of an .ncu-rep (needs -lineinfo and
--import-source on).  python tools/ncu_lines.py <file.ncu-rep> [nlines] [--sass lineNo]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
h = rows[hi]
ci = {c: i for i, c in enumerate(h)}
i_inst, i_samp, i_thr = ci["Instructions Executed"], ci["Warp Stall Sampling (All Samples)"], ci["Thread Instructions Executed"]
cur = None; agg = {}
sass_of = None
if "--sass" in sys.argv: sass_of = sys.argv[sys.argv.index("--sass") + 1]
for r in rows[hi + 1:]:
    if len(r) < len(h): continue
    if r[0].strip():                                   # a CUDA source line starts a group of SASS rows
        cur = (r[0], r[1].strip()); agg.setdefault(cur, [0.0, 0.0, 0.0])
        continue
    if cur is None: continue
    try:
        a = agg[cur]; a[0] += float(r[i_inst] or 0); a[1] += float(r[i_samp] or 0); a[2] += float(r[i_thr] or 0)
    except ValueError:
        continue
    if sass_of and cur[0] == sass_of:
        print("   %10s %6s  %s" % (r[i_inst], r[i_samp], r[3].strip()[:100]))
tot = sum(a[0] for a in agg.values()) or 1; tots = sum(a[1] for a in agg.values()) or 1
print("total warp instructions %.0f, samples %.0f" % (tot, tots))
for (ln, src), a in sorted(agg.items(), key=lambda x: -x[1][0])[:n]:
    print("%5.1f%% inst %5.1f%% samp  thr/inst %4.1f  L%-5s %s" % (100 * a[0] / tot, 100 * a[1] / tots, a[2] / max(1, a[0]), ln, src[:100]))
