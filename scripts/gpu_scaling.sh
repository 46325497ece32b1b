#!/bin/bash
# Runs on a multi-GPU box under `gpurun --gpus G`: NCCL parity of the sharded env at 2 ranks, then bench.py at N=1,2,4,..,G.
mkdir -p gpurun_out
G=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node 2 --master-port 29511 scripts/dist_parity.py > gpurun_out/dist_parity.log 2>&1; echo "dist parity exit $?" >> gpurun_out/dist_parity.log
tail -4 gpurun_out/dist_parity.log
: > gpurun_out/scaling.jsonl
for n in 1 2 4 8; do
  [ $n -gt $G ] && break
  if [ $n -eq 1 ]; then CMD="python bench.py"; else CMD="$TR --nproc-per-node $n --master-port $((29520+n)) bench.py"; fi
  timeout 600 $CMD --gpus $n --steps 10 --warmup 3 --no-cpu-baseline --no-small-field > gpurun_out/scale_n$n.log 2> gpurun_out/scale_n$n.err
  echo "n=$n exit $?"; grep '^{' gpurun_out/scale_n$n.log >> gpurun_out/scaling.jsonl
done
python - <<'PY'
import json
rows=[json.loads(l) for l in open('gpurun_out/scaling.jsonl')]
base=rows[0]['value'] if rows else None
for r in rows:
    print(f"n_gpus={r['n_gpus']} ms/step={r['ms_per_step']:.3f} value={r['value']:.4e} e2e={r['e2e']['value']:.4e} eff={r['value']/(base*r['n_gpus']):.3f} clocks={r['clocks']}")
PY
