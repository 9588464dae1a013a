#!/bin/bash
# Developer tool: build an A/B variant of libhuffb200.so with extra -D flags.
#   tools/build_variant.sh <tag> [-DNAME=VALUE ...]   ->  tools/variants/libhuffb200_<tag>.so  (select with HZ_LIB=<path>)
set -e
TAG=$1; shift
ROOT=$(cd "$(dirname "$0")/.." && pwd)
PKG="$ROOT/data-compression-implementing-gpu-driven-huffman-encoding-in-java_b200"
OUT="$ROOT/tools/variants"; mkdir -p "$OUT/build_$TAG"
OBJS=""
for f in hz_api hz_hist hz_codebook hz_encode hz_decode hz_decode_fused hz_global hz_sha256 hz_synth hz_container; do
  src="$PKG/csrc/$f.cu"; x=""
  [ -f "$src" ] || { src="$PKG/csrc/$f.cpp"; x="-x cu"; }
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC,-O3,-pthread "$@" $x -c "$src" -o "$OUT/build_$TAG/$f.o" &
  OBJS="$OBJS $OUT/build_$TAG/$f.o"
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o "$OUT/libhuffb200_$TAG.so" $OBJS -Xcompiler -pthread -ldl
echo "$OUT/libhuffb200_$TAG.so"
