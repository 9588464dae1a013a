"""Developer tool: pinned-memory host<->device copy bandwidth of the box, the bound of bench.py's e2e number.
One GPU:      python tools/pcie_peak.py [MiB]
G GPUs at once (how much the HOST sustains when every rank copies, the ceiling of the multi-GPU e2e curve):
              python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 tools/pcie_peak.py [MiB] --json out.json
Every rank copies H2D alone, D2H alone and both at once between barriers; the figures are the per-GPU GB/s of the
SLOWEST rank (that is what bounds a step timed as the max over ranks) and the sum over ranks."""
import json
import os
import sys

import torch

args = [a for a in sys.argv[1:] if not a.startswith("--")]
n = (int(args[0]) if args else 2048) << 20
out_json = sys.argv[sys.argv.index("--json") + 1] if "--json" in sys.argv else None
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
try:                                        # same NUMA binding as bench.py
    import importlib
    import __graft_entry__ as ge
    importlib.import_module(ge.load_package().__name__ + ".parallel").bind_to_gpu_numa(local)
except Exception:
    pass
h_a = torch.empty(n, dtype=torch.uint8).pin_memory()
h_b = torch.empty(n, dtype=torch.uint8).pin_memory()
d_a = torch.empty(n, dtype=torch.uint8, device="cuda")
d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def barrier():
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
        torch.cuda.synchronize()


def timed(fn, reps=5):
    best = 1e9
    for _ in range(reps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        torch.cuda.synchronize()
        e1.record(); e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    t = torch.tensor([best], dtype=torch.float64, device="cuda")
    if dist:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)            # the slowest rank bounds a max-over-ranks step
    return float(t.item())


def h2d():
    with torch.cuda.stream(s1):
        d_a.copy_(h_a, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        h_b.copy_(d_b, non_blocking=True)


def both():
    h2d(); d2h()


res = {}
for key, name, fn in (("h2d_GBps", "H2D", h2d), ("d2h_GBps", "D2H", d2h),
                      ("duplex_GBps_per_direction", "H2D + D2H at once (per direction)", both)):
    ms = timed(fn)
    res[key] = n / ms / 1e6
    if rank == 0:
        print("%d GPU(s) at once  %-36s %7.2f ms  %6.1f GB/s per GPU  %7.1f GB/s host total" % (world, name, ms, n / ms / 1e6, world * n / ms / 1e6))
if rank == 0 and out_json:
    try:
        with open(out_json) as f:
            allres = json.load(f)
    except Exception:
        allres = {}
    res["gpus"] = world; res["mib_per_copy"] = n >> 20
    res["note"] = "per-GPU GB/s of the slowest rank, all ranks copying between barriers (tools/pcie_peak.py)"
    allres[str(world)] = res
    with open(out_json, "w") as f:
        json.dump(allres, f, indent=1)
if dist:
    dist.destroy_process_group()
