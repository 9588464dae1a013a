"""Developer tool: pinned-memory PCIe copy bandwidth of the box (H2D alone, D2H alone, both at once), the
bound of bench.py's e2e number.  python tools/pcie_peak.py [MiB]"""
import sys
import torch

n = (int(sys.argv[1]) if len(sys.argv) > 1 else 2048) << 20
h_a = torch.empty(n, dtype=torch.uint8).pin_memory()
h_b = torch.empty(n, dtype=torch.uint8).pin_memory()
d_a = torch.empty(n, dtype=torch.uint8, device="cuda")
d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=5):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        torch.cuda.synchronize()
        e1.record(); e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def h2d():
    with torch.cuda.stream(s1):
        d_a.copy_(h_a, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        h_b.copy_(d_b, non_blocking=True)


def both():
    h2d(); d2h()


for name, fn, mult in (("H2D", h2d, 1), ("D2H", d2h, 1), ("H2D + D2H at once (per direction)", both, 1)):
    ms = timed(fn)
    print("%-36s %7.2f ms  %6.1f GB/s" % (name, ms, mult * n / ms / 1e6))
