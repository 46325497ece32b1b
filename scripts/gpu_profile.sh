#!/bin/bash
# Runs on the GPU box under gpurun: ncu launch list of one bench step + ncu --set full of the two splat kernels.
# Each ncu run is preceded by the identical plain command (must exit 0).  Everything lands in gpurun_out/.
mkdir -p gpurun_out
B=${PROF_B:-4096}
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-small-field --no-culled"
$BENCH > gpurun_out/prof_bench_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $BENCH > gpurun_out/prof_bench_ncu.log 2>&1
echo "launch list exit $?"
for what in fwd bwd; do
  CMD="python scripts/prof_splat.py --what $what --impl 2 --B $B --iters 2"
  $CMD > gpurun_out/prof_${what}_plain.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:splat_${what}_tc -s 2 -c 1 -f -o gpurun_out/prof_${what}_tc $CMD > gpurun_out/prof_${what}_ncu.log 2>&1
  echo "$what full exit $?"; cat gpurun_out/prof_${what}_plain.log
done
ls -la gpurun_out
