#!/bin/bash
# Runs on the GPU box (under gpurun): GPU parity tests, the bench, the ncu launch list and one
# `ncu --set full` capture per hot kernel.  Outputs land in gpurun_out/.
#   tools/gpu_profile.sh [tag] [kernel-regex ...]
set -u
TAG=${1:-run}; shift || true
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.txt 2>&1
if [ "${SKIP_TESTS:-0}" != "1" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
  tail -3 $OUT/pytest_gpu.log
fi
timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
cat $OUT/bench.json | head -c 6000; tail -5 $OUT/bench.err
if [ "${SKIP_REF:-0}" != "1" ]; then
  timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "ref rc=$?"
  cat $OUT/bench_ref.json
fi
if [ "${SKIP_NCU:-0}" != "1" ]; then
  CMD="python bench.py --steps 2 --warmup 3 --size-mib ${NCU_MIB:-1024} --no-e2e --no-cpu --no-strong"
  $CMD > $OUT/ncu_plain.log 2>&1 &&
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv $CMD > $OUT/ncu_launches.log 2>&1
  echo "ncu launches rc=$?"
  for K in "$@"; do
    timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -f -o $OUT/prof_$K $CMD > $OUT/ncu_$K.log 2>&1
    echo "ncu $K rc=$?"
  done
fi
ls -la $OUT
