#!/usr/bin/env python
"""Runs an UNMODIFIED reference script with the drop-in modules shadowing the reference's own:

    python /path/to/mofo-b200/dropin/launch.py /path/to/MOFO/run_mae_pretraining_BB.py --batch_size 32 ...
    python -m torch.distributed.run --nproc-per-node 8 /path/to/mofo-b200/dropin/launch.py /path/to/MOFO/run_mae_pretraining_BB.py ...

``python script.py`` puts the script's directory FIRST on sys.path, ahead of PYTHONPATH, so the reference's own
masking_generator / modeling_pretrain / engine_for_pretraining / utils / optim_factory would win.  This launcher orders
sys.path as [dropin, repo root, reference checkout, ...] and then executes the script as __main__."""
import os
import runpy
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    if len(sys.argv) < 2:
        sys.exit(__doc__)
    script = os.path.abspath(sys.argv[1])
    ref_dir = os.path.dirname(script)
    os.environ.setdefault("MOFO_REFERENCE_DIR", ref_dir)
    for p in (HERE, os.path.dirname(HERE), ref_dir):
        while p in sys.path:
            sys.path.remove(p)
    sys.path[0:0] = [HERE, os.path.dirname(HERE), ref_dir]
    sys.argv = [script] + sys.argv[2:]
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
