#!/bin/bash
# developer experiment: rebuild the encoder with ENC_EXP=n (parts of the kernel disabled; output is
# wrong on purpose) and time it, to attribute kernel time to its phases
cd data-compression-implementing-gpu-driven-huffman-encoding-in-java_b200
for e in "$@"; do
  rm -f build/hz_encode.o; make -s EXTRA=-DENC_EXP=$e libhuffb200.so > /dev/null 2>&1
  echo "EXP=$e"; (cd ..; python tools/quickbench.py 1024 ${H:-4} | grep -E "^  encode|roundtrip")
done
rm -f build/hz_encode.o; make -s libhuffb200.so > /dev/null 2>&1
