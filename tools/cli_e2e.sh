#!/bin/bash
# End-to-end CLI number: `datacomp compress` / `decompress` of a 1 GiB Zipf file on /dev/shm
# (file I/O, SHA-256 of every chunk and the container included), beside the CPU port of the reference path
# (oracle compress()/decompress(), 8 worker threads, 256 MiB sample, in memory).
set -e
D=/dev/shm/hzcli; mkdir -p $D
python - <<'PY'
import sys, os, time
sys.path.insert(0, "tests")
import numpy as np, datasets, orc
q = datasets.zipf_qtable(4)
with open("/dev/shm/hzcli/in.bin", "wb") as f:
    for o in range(0, 1 << 30, 64 << 20):
        f.write(datasets.synth_host(64 << 20, 0x5EED0001, q, o).tobytes())
s = np.fromfile("/dev/shm/hzcli/in.bin", dtype=np.uint8, count=256 << 20)
t0 = time.perf_counter(); blob = orc.compress(s, 16 << 20, literal=True); t1 = time.perf_counter()
back = orc.decompress(blob, literal=True); t2 = time.perf_counter()
assert np.array_equal(np.frombuffer(back, dtype=np.uint8) if not isinstance(back, np.ndarray) else back, s)
print("CPU port (8 threads, 256 MiB, in memory, SHA-256 included): compress %.3f s = %.3f GB/s, decompress %.3f s = %.3f GB/s"
      % (t1 - t0, s.size / (t1 - t0) / 1e9, t2 - t1, s.size / (t2 - t1) / 1e9))
PY
B=data-compression-implementing-gpu-driven-huffman-encoding-in-java_b200/datacomp
$B c $D/in.bin $D/out.dcz 16 > /dev/null    # warm-up (context creation, first-touch)
for i in 1 2; do
  t0=$(date +%s.%N); $B c $D/in.bin $D/out.dcz 16 > $D/c.log; t1=$(date +%s.%N)
  $B d $D/out.dcz $D/back.bin > $D/d.log; t2=$(date +%s.%N)
  python -c "print('datacomp compress 1 GiB: %.3f s wall = %.2f GB/s; decompress: %.3f s wall = %.2f GB/s (process start, CUDA context, file I/O on /dev/shm, SHA-256, container included)' % ($t1-$t0, 1.0737/($t1-$t0), $t2-$t1, 1.0737/($t2-$t1)))"
  grep -iE "throughput|ratio" $D/c.log $D/d.log || true
done
cmp $D/in.bin $D/back.bin && echo "round trip identical"; ls -l $D; rm -rf $D
