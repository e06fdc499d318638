"""Stage an UNMODIFIED copy of the reference's MonoDETR Python tree for the training-step harness.

    python tools/stage_reference.py        # needs /root/reference; writes baseline/_ref/MonoDETR/

Why: SURVEY.md 8 row f1 measures "the unmodified reference model with this repo's op swapped in".
/root/reference does not exist on the GPU box, so the model code has to travel with the repo
snapshot.  baseline/_ref/ is git-ignored (never committed, never part of this repo's source) but not
gpurun-ignored -- the same place the base contract uses for the reference install.  Only .py / .yaml
files of lib/, utils/ and configs/ are copied; the compiled-extension sources (ops/src), checkpoints,
image sets and data are not.  Nothing is edited: every incompatibility with torch >= 2 is handled by
import shims in tools/train_step_bench.py.
"""
import os
import shutil
import sys

SRC = "/root/reference/MonoDETR"
DST = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref", "MonoDETR")


def main():
    if not os.path.isdir(SRC):
        print(f"{SRC} not present; keeping whatever is staged in {DST}")
        return 0
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    n = 0
    for top in ("lib", "utils", "configs"):
        for root, dirs, files in os.walk(os.path.join(SRC, top)):
            dirs[:] = [d for d in dirs if d not in ("__pycache__", "src", "kitti_eval_python")]
            for f in files:
                if f.endswith((".py", ".yaml")):
                    rel = os.path.relpath(os.path.join(root, f), SRC)
                    os.makedirs(os.path.dirname(os.path.join(DST, rel)), exist_ok=True)
                    shutil.copy2(os.path.join(root, f), os.path.join(DST, rel))
                    n += 1
    print(f"staged {n} files into {DST}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
