#!/bin/bash
# Runs under `gpurun --gpus G`: NCCL parity of the sharded env, bench.py at G ranks (weak headline + strong block), optionally the sweep.
mkdir -p gpurun_out
G=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node 2 --master-port 29511 scripts/dist_parity.py > gpurun_out/dist_parity.log 2>&1; echo "dist parity exit $?" >> gpurun_out/dist_parity.log
tail -6 gpurun_out/dist_parity.log
for n in ${SCALE_NS:-$G}; do
  if [ $n -eq 1 ]; then CMD="python bench.py"; else CMD="$TR --nproc-per-node $n --master-port $((29520+n)) bench.py"; fi
  timeout 900 $CMD --gpus $n --steps ${STEPS:-10} --warmup 3 --no-cpu-baseline --no-small-field --no-culled --no-gpu-eager > gpurun_out/scale_n$n.log 2> gpurun_out/scale_n$n.err
  echo "bench n=$n exit $?"; tail -3 gpurun_out/scale_n$n.err
done
if [ -n "$SWEEP" ]; then
  timeout 1500 $TR --nproc-per-node $G --master-port 29540 scripts/sweep.py --steps 3 > gpurun_out/sweep_n$G.jsonl 2> gpurun_out/sweep_n$G.err
  echo "sweep n=$G exit $?"; tail -3 gpurun_out/sweep_n$G.err
fi
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/scale_n*.log')):
    for l in open(f):
        if l.startswith('{'):
            r = json.loads(l)
            s = r.get('strong_scaling') or {}
            print(f"n_gpus={r['n_gpus']} ms/step={r['ms_per_step']:.3f} value={r['value']:.4e} e2e={r['e2e']['value']:.4e} e2e_ms={r['e2e']['ms_per_step']:.3f} "
                  f"link={r['e2e'].get('host_link')} strong={s.get('ms_per_step')} {s.get('value')} uncached={(r.get('uncached') or {}).get('ms_per_step')}")
PY
