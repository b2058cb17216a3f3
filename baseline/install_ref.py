#!/usr/bin/env python
"""Install the UNMODIFIED reference into baseline/_ref (git-ignored, but it travels to the GPU box with gpurun).

The reference has neither setup.py nor pyproject.toml, so
    python -m pip install --no-index --no-build-isolation --find-links /opt/wheelhouse --target baseline/_ref /root/reference
fails with "Directory '/root/reference' is not installable" (recorded in DESIGN.md section 9).  What pip --target would
do for a pure-Python project - place its importable modules in the target directory - is done here directly: the model
package (torch-only imports) plus checkpoint.py / manager.py, byte for byte.  Nothing under baseline/_ref is ever
committed or imported by the product (conformer_b200/); only bench.py's reference arm / gpu_eager_baseline and the
tests' fixtures use it, and they fall back to the oracle port when it is absent.
"""
import os
import shutil
import sys

SRC = "/root/reference"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
ITEMS = ["model", "checkpoint.py", "manager.py"]


def main() -> int:
    if not os.path.isdir(SRC):
        print(f"{SRC} not present (GPU box?): nothing to install", file=sys.stderr)
        return 0
    os.makedirs(DST, exist_ok=True)
    for item in ITEMS:
        s, d = os.path.join(SRC, item), os.path.join(DST, item)
        if os.path.isdir(s):
            shutil.rmtree(d, ignore_errors=True)
            shutil.copytree(s, d, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
        else:
            shutil.copyfile(s, d)
    print(f"installed {ITEMS} from {SRC} into {DST}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
