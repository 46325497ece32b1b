#!/bin/bash
# A/B of whole-step kernel times per library build: LIBS="libhelio_sm100.so libhelio_v6.so" bash scripts/gpu_ab_step.sh
mkdir -p gpurun_out
: > gpurun_out/ab_step.log
for rep in 1 2; do
  for lib in ${LIBS:-libhelio_sm100.so}; do
    for shape in "2000 256 4096" "500 128 16384" "5000 64 4096"; do
      echo -n "$lib: " >> gpurun_out/ab_step.log
      HELIO_LIB_PATH=$PWD/doodle_b200/$lib timeout 200 python scripts/ab_step.py $shape 2>&1 | tail -1 >> gpurun_out/ab_step.log
    done
  done
done
cat gpurun_out/ab_step.log
