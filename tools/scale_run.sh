#!/bin/bash
# Runs on the GPU box under `gpurun --gpus N`: the driver's bench command line at N GPUs (weak scaling + the
# strong-scaling section + the sharded single-file parity check), the same in global-codebook mode, and the
# concurrent host<->device copy ceiling.  Outputs: gpurun_out/scale/.
#   tools/scale_run.sh N [steps]
set -u
N=$1; STEPS=${2:-10}
OUT=gpurun_out/scale; mkdir -p $OUT
RUN="python"
[ "$N" -gt 1 ] && RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 900 $RUN bench.py --gpus $N --steps $STEPS --warmup 3 > $OUT/n$N.json 2> $OUT/n$N.err; echo "bench rc=$?"
timeout 600 $RUN bench.py --gpus $N --steps $STEPS --warmup 3 --codebook global --no-e2e --no-cpu > $OUT/n${N}_global.json 2> $OUT/n${N}_global.err; echo "global rc=$?"
timeout 300 $RUN tools/pcie_peak.py 1024 --json $OUT/pcie.json > $OUT/pcie_n$N.txt 2>&1; echo "pcie rc=$?"
python - <<PY
import json
for f in ("$OUT/n$N.json", "$OUT/n${N}_global.json"):
    try:
        d = json.load(open(f))
        print(f, "value %.1f ms/step %.3f" % (d["value"], d["ms_per_step"]), "e2e", d["e2e"] and round(d["e2e"]["value"], 1),
              "strong", d["strong_scaling"] and round(d["strong_scaling"]["value"], 1), "sharded", d["sharded"] and d["sharded"]["sharded_parity"],
              "roofline", round(d["roofline"]["frac"], 3))
    except Exception as e:
        print(f, "ERR", e)
PY
cat $OUT/pcie_n$N.txt | tail -3; tail -2 $OUT/n$N.err
