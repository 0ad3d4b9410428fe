#!/bin/bash
# Development sweep: forward / adjoint timings over batch sizes for each side-by-side build variant of libdwdf
# (csrc/Makefile: TILE / STAGES). Usage: tools/sweep_chunks.sh [out_file]
out=${1:-gpurun_out/sweep_chunks.txt}
mkdir -p "$(dirname "$out")"
: > "$out"
for lib in libdwdf.so libdwdf_t32s2.so libdwdf_t16s3.so libdwdf_t16s4.so; do
  [ -f differentiable-wdfs_b200/$lib ] || continue
  echo "== $lib" >> "$out"
  for B in 256 1024 4096 8192 16384 32768 65536; do
    DWDF_LIBRARY=$PWD/differentiable-wdfs_b200/$lib timeout 300 python tools/kbench.py --B $B --iters 7 >> "$out" 2>&1
  done
done
cat "$out"
