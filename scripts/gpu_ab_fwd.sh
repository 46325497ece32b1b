#!/bin/bash
# A/B of forward-splat builds on one box: LIBS="libhelio_sm100.so libhelio_v1.so ..." bash scripts/gpu_ab_fwd.sh
# (variants are built with `make -C doodle_b200/csrc dbg DBGFLAGS=... DBGOUT=../libhelio_vN.so`)
mkdir -p gpurun_out
: > gpurun_out/ab_fwd.log
LIBS=${LIBS:-"libhelio_sm100.so"}
for rep in 1 2; do
  for lib in $LIBS; do
    for shape in "--N 2000 --R 256 --B 4096" "--N 500 --R 128 --B 16384" "--N 5000 --R 64 --B 4096" "--N 50 --R 128 --B 16384" "--N 500 --R 512 --B 1024"; do
      echo -n "$lib $shape: " >> gpurun_out/ab_fwd.log
      HELIO_LIB_PATH=$PWD/doodle_b200/$lib timeout 120 python scripts/prof_splat.py --what ${WHAT:-fwd} --impl 2 $shape --iters 3 ${EXTRA} >> gpurun_out/ab_fwd.log 2>&1
    done
  done
done
cat gpurun_out/ab_fwd.log
