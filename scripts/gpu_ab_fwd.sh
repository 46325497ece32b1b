#!/bin/bash
# A/B of forward-splat builds on one box: timing (prof_splat.py) per library, per-role cycle accounting with the stats build.
mkdir -p gpurun_out
: > gpurun_out/ab_fwd.log
for lib in libhelio_sm100.so libhelio_pf1.so; do
  for rep in 1 2; do
    for shape in "--N 2000 --R 256 --B 4096" "--N 500 --R 128 --B 16384" "--N 5000 --R 64 --B 4096" "--N 50 --R 128 --B 16384"; do
      echo -n "$lib $shape: " >> gpurun_out/ab_fwd.log
      HELIO_LIB_PATH=$PWD/doodle_b200/$lib python scripts/prof_splat.py --what fwd --impl 2 $shape --iters 3 >> gpurun_out/ab_fwd.log 2>&1
    done
    echo -n "$lib f16x3 N2000 R256: " >> gpurun_out/ab_fwd.log
    HELIO_FWD_PREC=1 HELIO_LIB_PATH=$PWD/doodle_b200/$lib python scripts/prof_splat.py --what fwd --impl 2 --N 2000 --R 256 --B 4096 --iters 3 >> gpurun_out/ab_fwd.log 2>&1
  done
done
HELIO_LIB_PATH=$PWD/doodle_b200/libhelio_dbg.so python scripts/tc_stats.py --what fwd > gpurun_out/tc_stats.log 2>&1
HELIO_LIB_PATH=$PWD/doodle_b200/libhelio_dbg.so python scripts/tc_stats.py --what fwd --prec 1 >> gpurun_out/tc_stats.log 2>&1
cat gpurun_out/ab_fwd.log; cat gpurun_out/tc_stats.log
