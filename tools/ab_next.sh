#!/bin/bash
# A/B of the compile-time variants prepared at the end of round 1:
#   tools/ab_next.sh build     here (nvcc cross-compiles): tools/variants/libhuffb200_<tag>.so for every variant
#   gpurun --timeout 900 -- tools/ab_next.sh run      on the GPU box: parity suite + per-kernel times for each variant
# Results: gpurun_out/ab_next/<tag>.txt (last lines of pytest, then tools/kprof.py on the 4 GiB bench workload).
set -u
ROOT=$(cd "$(dirname "$0")/.." && pwd)
cd "$ROOT"
declare -A VAR=(
  [enc1]="-DHZ_ENC_IMAD_EXTRACT=1"
  [enc2]="-DHZ_ENC_IMAD_EXTRACT=2"
  [enc3]="-DHZ_ENC_IMAD_EXTRACT=3"
  [decfma]="-DHZ_DEC_FMA_SHIFTS=1"
  [decfma2]="-DHZ_DEC_FMA_SHIFTS=2"
  [all]="-DHZ_ENC_IMAD_EXTRACT=2 -DHZ_DEC_FMA_SHIFTS=2"
)
case "${1:-}" in
  build)
    for tag in "${!VAR[@]}"; do tools/build_variant.sh "$tag" ${VAR[$tag]} | tail -1; done ;;
  run)
    OUT=gpurun_out/ab_next; mkdir -p $OUT
    REPS=${REPS:-5} timeout 120 python tools/kprof.py 4096 4:16384 > $OUT/base.txt 2>&1
    for tag in "${!VAR[@]}"; do
      lib=$ROOT/tools/variants/libhuffb200_$tag.so
      [ -f "$lib" ] || { echo "missing $lib (run: tools/ab_next.sh build)"; continue; }
      HZ_LIB=$lib timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > $OUT/$tag.txt
      HZ_LIB=$lib REPS=${REPS:-5} timeout 120 python tools/kprof.py 4096 4:16384 >> $OUT/$tag.txt 2>&1
    done
    tail -n +1 $OUT/*.txt ;;
  *) echo "usage: $0 build | run"; exit 1 ;;
esac
