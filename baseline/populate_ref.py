"""Populate baseline/_ref/ with the UNMODIFIED reference sources of the path (ModelTC/TFMQ-DM: quant/, ddim/, linklink/,
stable-diffusion/ldm/ -- pure Python, no build step), so that `bench.py --impl reference` can run the reference's own
QuantModel on the GPU box's host cores.  baseline/_ref/ is git-ignored (the reference is not product source) but travels with
the gpurun snapshot.  Run in the build container:  python baseline/populate_ref.py
The reference is a script tree, not a package (`stable-diffusion/setup.py` covers `ldm` only, and `pip install` of it would
miss quant/ and ddim/), hence a copy instead of the pip recipe."""
import os
import shutil
import sys

SRC = os.environ.get("TFMQ_REFERENCE", "/root/reference")
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def main():
    if not os.path.isdir(SRC):
        print(f"{SRC} not found: nothing copied (the reference arm then falls back to the oracle port)")
        return 1
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    keep = lambda d, names: [n for n in names if not (n.endswith(".py") or os.path.isdir(os.path.join(d, n)))]  # noqa: E731
    for sub in ("quant", "ddim", "linklink", os.path.join("stable-diffusion", "ldm")):
        shutil.copytree(os.path.join(SRC, sub), os.path.join(DST, sub), ignore=keep)
    n = sum(len([f for f in fs if f.endswith(".py")]) for _, _, fs in os.walk(DST))
    print(f"copied {n} reference .py files to {DST}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
