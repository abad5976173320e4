"""Places the UNMODIFIED reference package under baseline/_ref/ (git-ignored; it travels to
the GPU box with the snapshot) so `bench.py --impl reference` can time the reference's own
public API there.

The contract's `pip install --target baseline/_ref /root/reference` needs the reference's
build backend (hatchling), which is not in the offline wheelhouse; the wheel it would build
is pure Python (pyproject.toml: packages = ["src/scikits"]), i.e. exactly the tree copied
here.  Product code never imports baseline/_ref (tests/test_host_probes.py checks it).

    python baseline/install_ref.py [/root/reference]
"""
import os
import shutil
import sys


def main(ref="/root/reference"):
    here = os.path.dirname(os.path.abspath(__file__))
    dst = os.path.join(here, "_ref")
    src = os.path.join(ref, "src", "scikits")
    if not os.path.isdir(src):
        print(f"reference not found at {src}; baseline/_ref left as is")
        return 1
    shutil.rmtree(dst, ignore_errors=True)
    os.makedirs(dst)
    shutil.copytree(src, os.path.join(dst, "scikits"), ignore=shutil.ignore_patterns("__pycache__"))
    print("installed", os.path.join(dst, "scikits"))
    return 0


if __name__ == "__main__":
    sys.exit(main(*sys.argv[1:]))
