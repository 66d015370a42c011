"""Launcher that runs an UNMODIFIED reference entry point on the sm_100a hot path:

    LAYOUTDETR_REFERENCE=/path/to/LayoutDETR python -m layoutdetr_b200.dropin.run train.py --gpus=8 --batch=16 ...

`python train.py` would put the checkout at sys.path[0], ahead of any PYTHONPATH overlay, so this launcher inserts the
overlay directory first, the checkout second, and executes the script with runpy (spawned workers inherit sys.path)."""
import os
import runpy
import sys


def main():
    ref = os.environ.get("LAYOUTDETR_REFERENCE", "/root/reference")
    if len(sys.argv) < 2:
        raise SystemExit("usage: python -m layoutdetr_b200.dropin.run <script.py> [args...]")
    overlay = os.path.dirname(os.path.abspath(__file__))
    script = sys.argv[1]
    if not os.path.isabs(script):
        script = os.path.join(ref, script)
    sys.argv = [script] + sys.argv[2:]
    sys.path[:0] = [overlay, ref]
    os.environ.setdefault("NCCL_P2P_LEVEL", "NVL")      # data-parallel workers (train.py spawns them): peer traffic over NVLink only
    os.chdir(ref)                       # the reference opens configs/med_config.json and pretrained/ by relative path
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
