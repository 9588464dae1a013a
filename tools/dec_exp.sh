#!/bin/bash
# developer experiment: rebuild the decoder with -D<define> variants and time the kernels
cd data-compression-implementing-gpu-driven-huffman-encoding-in-java_b200
for e in "$@"; do
  rm -f build/hz_decode.o; make -s EXTRA="-D$e" libhuffb200.so > /dev/null 2>&1
  echo "== $e"; (cd ..; python tools/quickbench.py 1024 ${H:-4} | grep -E "^  dec_sync|^  dec_write|^  dec_fix|roundtrip")
done
rm -f build/hz_decode.o; make -s libhuffb200.so > /dev/null 2>&1
