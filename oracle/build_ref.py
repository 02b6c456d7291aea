"""Recipe: materialise the reference's own Python modules under oracle/_ref/ (git-ignored, shipped to the GPU box).

The reference (mrafidashti/neuradar, a nerfstudio fork) is pure Python; its fp32 torch path
(`implementation="torch"`) is the parity target and the CPU baseline of BASELINE.md section 4.  /root/reference
does not exist on the GPU box, so this script copies the `nerfstudio` package's *.py files (2.7 MB, nothing else)
to oracle/_ref/nerfstudio where `oracle.ref_shim` finds them.  Nothing under oracle/_ref/ is tracked by git, and
nothing in the product package imports it: it is test / baseline infrastructure only.

  python oracle/build_ref.py [--reference /root/reference]
"""
import argparse
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")


def build(reference: str = "/root/reference", quiet: bool = False) -> bool:
    src = os.path.join(reference, "nerfstudio")
    if not os.path.isdir(src):
        if not quiet:
            print(f"oracle/build_ref: no reference checkout at {reference}; keeping {DEST} as it is")
        return False
    dst = os.path.join(DEST, "nerfstudio")
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    n = 0
    for root, _dirs, files in os.walk(src):
        rel = os.path.relpath(root, src)
        for f in files:
            if f.endswith(".py"):
                os.makedirs(os.path.join(dst, rel), exist_ok=True)
                shutil.copyfile(os.path.join(root, f), os.path.join(dst, rel, f))
                n += 1
    with open(os.path.join(DEST, "SOURCE.txt"), "w") as fh:
        fh.write(f"copied {n} *.py files of {src} (unmodified) by oracle/build_ref.py\n")
    if not quiet:
        print(f"oracle/build_ref: {n} reference modules -> {dst}")
    return True


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default=os.environ.get("NEURADAR_REFERENCE", "/root/reference"))
    args = ap.parse_args()
    sys.exit(0 if build(args.reference) else 1)
