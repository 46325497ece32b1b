#!/bin/bash
# A/B of the backward operand formats inside the fused step (3xTF32 vs f16x3 K=64): parity subset under each, then the bench.
mkdir -p gpurun_out
for pr in 1 0; do
  HELIO_BWD_PREC=$pr timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "env_step_matches_reference or fused_step_equals or transparent_graph or options_against_oracle or culled_step or host_action" > gpurun_out/pytest_bwdprec$pr.log 2>&1
  echo "prec=$pr pytest: $(tail -1 gpurun_out/pytest_bwdprec$pr.log)"
  HELIO_BWD_PREC=$pr timeout 300 python bench.py --steps 10 --no-culled --no-gpu-eager --no-cpu-baseline --no-small-field > gpurun_out/bench_bwdprec$pr.log 2> gpurun_out/bench_bwdprec$pr.err
  python - <<PY
import json
for l in open("gpurun_out/bench_bwdprec$pr.log"):
    if l.startswith("{"):
        r=json.loads(l); print("prec=$pr", r["ms_per_step"], r["roofline"]["kernels_ms"], r["roofline"]["sm_mhz_held_in_kernel"]["splat_bwd"], "uncached", r["uncached"]["ms_per_step"])
PY
done
