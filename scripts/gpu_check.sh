#!/bin/bash
# Runs on the GPU box under gpurun: parity tests, smoke, a short bench.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/box.log 2>&1
(nproc; free -g | head -2) >> gpurun_out/box.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q ${PYTEST_ARGS:--x} --timeout=900 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 900 python bench.py --steps ${STEPS:-5} --warmup 3 --cpu-seconds 15 ${BENCH_ARGS} > gpurun_out/bench.log 2> gpurun_out/bench.err
echo "bench exit $?" >> gpurun_out/bench.err
tail -25 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log | tail -3; cat gpurun_out/bench.log; tail -5 gpurun_out/bench.err
