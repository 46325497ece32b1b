#!/bin/bash
# Acceptance harness: run the reference's OWN trainer, unmodified, on top of the sm_100a HelioField / HelioEnv.
#
#   scripts/run_reference_trainer.sh /path/to/DOODLE [extra train_with_env.py arguments]
#
# dropin/ shadows the two hot-path modules the trainer imports by name (test_environment, newenv_rl_test_multi_error);
# scripts/run_with_dropin.py puts it in front of the trainer's own directory; scripts/ref_stubs/ supplies stand-ins for third-party packages that are not installed in this image (adamp, mlflow,
# plotly, matplotlib, gymnasium -- see scripts/ref_stubs/README.md).  Needs a cc-10.x GPU and a checkout of the reference
# (BASELINE.json configs[2]: LSTM policy, N=50, 128x128, B=256, alignment pretrain + warm-up schedule).  The trainer
# needs --num_batches >= 2 (it steps the optimiser on i % (num_batches - 1), train_with_env.py:383) and a batch of
# at least 60 suns (its test env slices the first 60, :259,:275).
set -e
REF=${1:?usage: run_reference_trainer.sh /path/to/DOODLE [args...]}
shift || true
REPO="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
[ -f "$REF/train_with_env.py" ] || { echo "no train_with_env.py under $REF" >&2; exit 2; }
# (python puts a script's own directory first on sys.path, ahead of PYTHONPATH, so the launcher arranges the path itself)
exec python "$REPO/scripts/run_with_dropin.py" "$REF/train_with_env.py" --device cuda --num_batches 2 --batch_size 256 --steps 5 --T 4 --k 4 "$@"
