"""Stage the reference's own pretraining-path modules, UNMODIFIED, under the git-ignored baseline/_ref/ so they travel to
the GPU box with the working-tree snapshot (they are never committed; .gitignore lists baseline/_ref/).

    python baseline/setup_ref.py            # needs /root/reference (build container only)

Called by __graft_entry__.build().  On the GPU box /root/reference does not exist and the staged copies are used as-is.
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("MOFO_REFERENCE", "/root/reference")
REF_DST = os.path.join(HERE, "_ref")
FILES = ["masking_generator.py", "modeling_pretrain.py", "modeling_finetune.py", "engine_for_pretraining.py", "utils.py",
         "optim_factory.py", "engine_for_finetuning.py", "mixup.py", os.path.join("scripts", "motion_sts.py")]


def stage(verbose=False):
    """Returns the list of staged files ([] when the reference tree is not present)."""
    if not os.path.isdir(REF_SRC):
        return []
    os.makedirs(REF_DST, exist_ok=True)
    done = []
    for f in FILES:
        s, d = os.path.join(REF_SRC, f), os.path.join(REF_DST, os.path.basename(f))
        if not os.path.exists(s):
            continue
        if not os.path.exists(d) or os.path.getmtime(d) < os.path.getmtime(s) or os.path.getsize(d) != os.path.getsize(s):
            shutil.copyfile(s, d)
        done.append(d)
    if verbose:
        print(f"staged {len(done)} reference modules into {REF_DST}")
    return done


def available():
    return all(os.path.exists(os.path.join(REF_DST, f)) for f in FILES[:6])


if __name__ == "__main__":
    stage(verbose=True)
    sys.exit(0 if available() else 1)
