"""Launcher that puts the B200 ops in front of an UNMODIFIED reference script:

    PYTHONPATH=point-cloud-compression_b200 python -m pcc_b200.run /path/to/reference/compress.py <script args...>

The reference scripts parse argv and run at import time and import pytorch3d at the top (compress.py:8, eval.py:14-15),
so the shim must be installed before the script starts; this does exactly that and then runs it with runpy.
"""
import os
import runpy
import sys

from .install import install, patch_reference_modules


def main():
    if len(sys.argv) < 2:
        raise SystemExit("usage: python -m pcc_b200.run <reference_script.py> [args...]")
    script = os.path.abspath(sys.argv[1])
    sys.argv = [script] + sys.argv[2:]
    sys.path.insert(0, os.path.dirname(script))
    install()
    import pn_kit  # noqa: F401  (binds pytorch3d names from the shim; then rebind its own FPS / gather helpers)
    patch_reference_modules()
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
