#!/usr/bin/env python3
"""Run a script of the reference UNMODIFIED on top of the sm_100a HelioField / HelioEnv.

    python scripts/run_with_dropin.py /path/to/DOODLE/train_with_env.py --device cuda --num_batches 2 ...

``python /path/to/DOODLE/train_with_env.py`` puts the script's own directory FIRST on sys.path, ahead of PYTHONPATH, so
the reference's own ``test_environment.py`` / ``newenv_rl_test_multi_error.py`` would win over ``dropin/``.  This launcher
runs the script with ``dropin/`` (module-name shims), the repo root (``doodle_b200``) and, for packages that are not
installed (``scripts/ref_stubs``: adamp, mlflow, plotly, matplotlib, gymnasium), in front of the script's directory.
Stand-ins are only used for packages that cannot be imported: an installed adamp / mlflow / ... is preferred.
"""
import importlib.util
import os
import runpy
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    if len(sys.argv) < 2:
        raise SystemExit(__doc__)
    script = os.path.abspath(sys.argv[1])
    ref_dir = os.path.dirname(script)
    stubs = os.path.join(REPO, "scripts", "ref_stubs")
    missing = [m for m in ("adamp", "mlflow", "plotly", "matplotlib", "gymnasium") if importlib.util.find_spec(m) is None]
    own = os.path.dirname(os.path.abspath(__file__))
    rest = [p for p in sys.path if os.path.abspath(p or ".") != own]
    sys.path[:] = [os.path.join(REPO, "dropin"), REPO] + rest + [ref_dir] + ([stubs] if missing else [])
    if missing:
        print(f"[run_with_dropin] stand-ins for missing packages: {', '.join(missing)} (scripts/ref_stubs)", file=sys.stderr)
    sys.argv = [script] + sys.argv[2:]
    os.chdir(ref_dir)
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
