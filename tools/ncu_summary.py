#!/usr/bin/env python
"""Summarise an .ncu-rep (run where ncu is installed, no GPU needed): key raw metrics, warp-stall
breakdown and the hottest source lines.  python tools/ncu_summary.py <file.ncu-rep> [nlines]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
nlines = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "launch__grid_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__shared_mem_per_block_static", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_warps", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed"]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    for k in KEYS:
        if k in d:
            print("%-80s %-12s %s" % (k, units[hdr.index(k)], d[k]))
    st = [(float(d[h]), h) for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and d[h]]
    print("-- warp stalls per issue-active (top):")
    for v, h in sorted(st, reverse=True)[:8]:
        print("   %6.2f  %s" % (v, h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
# find header row
hi = next((i for i, r in enumerate(rows) if "Source" in r and any("Samples" in c for c in r)), None)
if hi is not None:
    h = rows[hi]
    ci_src = h.index("Source")
    ci_samp = next(i for i, c in enumerate(h) if c.startswith("# Samples") or c == "Warp Stall Sampling (All Samples)" or "Sampling (All" in c)
    ci_inst = next((i for i, c in enumerate(h) if c == "# Instructions Executed" or c.startswith("Instructions Executed")), None)
    body = []
    for r in rows[hi + 1:]:
        try:
            body.append((float(r[ci_samp] or 0), r))
        except (ValueError, IndexError):
            pass
    tot = sum(b[0] for b in body) or 1
    print("-- hottest lines (%s):" % h[ci_samp])
    for v, r in sorted(body, key=lambda x: -x[0])[:nlines]:
        print("   %5.1f%%  inst=%-10s %s" % (100 * v / tot, r[ci_inst] if ci_inst is not None else "", r[ci_src].strip()[:130]))
else:
    print("(no source page)")
