"""
Recipe that makes the REFERENCE ITSELF available where /root/reference does not exist (the GPU box).

    python oracle/ref_install.py            # run in the build container; __graft_entry__.build() calls it

It copies the reference's Python package (river-route v2.0.1, pure Python + numba: nothing to compile) from
/root/reference/river_route to oracle/_ref/river_route.  oracle/_ref/ is listed in .gitignore, so no reference
source ever enters this repository's history, but not in .gpurunignore, so the copy travels to the GPU box with
the snapshot exactly like the built .so files.  ``oracle/refarm.py`` imports it from there (with stub modules for
the I/O packages the image lacks) to time the reference's own numba kernels:
``bench.py --impl reference`` and the ``cpu_baseline`` leg.

Test / measurement infrastructure only: nothing under river_route_b200/ imports it.
"""
from __future__ import annotations

import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = '/root/reference/river_route'
DST = os.path.join(HERE, '_ref', 'river_route')


def install(verbose: bool = True) -> bool:
    """Copy the package when the reference tree is present; returns True when oracle/_ref holds a copy."""
    if os.path.isdir(SRC):
        if os.path.isdir(DST):
            shutil.rmtree(DST)
        os.makedirs(os.path.dirname(DST), exist_ok=True)
        shutil.copytree(SRC, DST, ignore=shutil.ignore_patterns('__pycache__', '*.pyc', '*.nbi', '*.nbc'))
        with open(os.path.join(HERE, '_ref', 'PROVENANCE.txt'), 'w') as f:
            f.write(f'copied from {SRC} by oracle/ref_install.py (unmodified)\n')
        if verbose:
            print(f'oracle/_ref: copied {SRC}')
    return os.path.isdir(DST)


if __name__ == '__main__':
    sys.exit(0 if install() else 1)
