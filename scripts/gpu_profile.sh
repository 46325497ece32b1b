#!/bin/bash
# Runs on the GPU box under gpurun: ncu launch list of one bench run + ncu --set full of the two splat kernels as HelioEnv.step launches them.
# Each ncu run is preceded by the identical plain command (must exit 0).  Everything lands in gpurun_out/.
mkdir -p gpurun_out
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-small-field --no-culled --no-gpu-eager --no-both-3xtf32"
$BENCH > gpurun_out/prof_bench_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $BENCH > gpurun_out/prof_bench_ncu.log 2>&1
echo "launch list exit $?"
CMD="python scripts/prof_step.py"
$CMD > gpurun_out/prof_step_plain.log 2>&1 || { echo "plain step failed"; cat gpurun_out/prof_step_plain.log; exit 1; }
for what in fwd bwd; do
  ncu --set full --clock-control none --import-source on -k regex:splat_${what}_tc -s 3 -c 1 -f -o gpurun_out/prof_${what}_tc $CMD > gpurun_out/prof_${what}_ncu.log 2>&1
  echo "$what full exit $?"
done
ls -la gpurun_out/*.ncu-rep
