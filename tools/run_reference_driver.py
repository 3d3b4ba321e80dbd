#!/usr/bin/env python
"""Run the reference's own driver, unmodified, against this repository's layers:

    python tools/run_reference_driver.py /path/to/CFUN/heart_main.py train --weights none --data ../data/ --stage beginning

Puts the repository root first on sys.path (so that `import model`, `utils`, `config`, `backbone`, `mask_branch` resolve to the
cfun_b200 shims, heart_main.py:15-17 / model.py:19-21), registers cfun_b200.nifti as `nibabel` when the real package is not
installed (heart_main.py:13), and executes the script as __main__ with the remaining arguments."""
import os
import runpy
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    if len(sys.argv) < 2:
        sys.exit(__doc__)
    script = os.path.abspath(sys.argv[1])
    sys.path.insert(0, ROOT)
    from cfun_b200 import nifti
    if nifti.install_as_nibabel():
        print("[cfun_b200] nibabel not installed: using cfun_b200.nifti", file=sys.stderr)
    sys.argv = [script] + sys.argv[2:]
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
