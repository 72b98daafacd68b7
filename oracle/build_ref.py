#!/usr/bin/env python
"""Compile the reference's own Cython modules for the hot path into oracle/_ref/.

TEST INFRASTRUCTURE ONLY.  Nothing under wann_b200/ may import oracle/.

The reference (TeoZosa/WaNN) is Python/Cython.  The parts of the hot path that do
NOT need TensorFlow -- board encoding (tools/utils.pyx), move-index tables
(tools/utils.pyx), legal-move enumeration (Breakthrough_Player/board_utils.pyx) and
the tree-side callers (monte_carlo_tree_search/*.pyx) -- compile with the Cython in
this image.  This script:

  1. copies the three source directories to a scratch dir under /tmp (the reference
     tree is read-only and its sources must never enter this repo),
  2. applies two one-line compatibility shims in the scratch copy only
       - tools/utils.pyx:16  `ctypedef np.int_t DTYPE_t`  (unused; np.int_t was
         removed from NumPy 2's pxd) is deleted,
       - `from bottleneck import argpartition` (board_utils.pyx:7) is satisfied at
         import time by oracle/ref_loader.py registering a stand-in module whose
         argpartition is numpy.argpartition (bottleneck is not installed),
  3. cythonizes (language=c++; tree_search_utils.pyx uses libcpp.bool) and copies
     ONLY the built extension modules + empty __init__.py files to oracle/_ref/.

oracle/_ref/ is git-ignored (binary artefacts) but travels to the GPU box.
The forward pass itself (TensorFlow 1.0.0 kernels) is NOT part of the reference
tree and cannot be built: see DESIGN.md "parity unpinned".
"""
import glob
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("WANN_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")

PKGS = ["tools", "Breakthrough_Player", "monte_carlo_tree_search"]
# breakthrough_player.pyx references an undefined name and does not compile under
# Cython 3; it is the game loop (out of scope) so it is skipped.
SKIP = {"breakthrough_player.pyx"}

SETUP = r'''
import sys, glob, os
from setuptools import setup, Extension
from Cython.Build import cythonize
import numpy
exts = []
for pkg in %(pkgs)r:
    for pyx in sorted(glob.glob(os.path.join(pkg, "*.pyx"))):
        if os.path.basename(pyx) in %(skip)r:
            continue
        mod = pyx[:-4].replace(os.sep, ".")
        exts.append(Extension(mod, [pyx], language="c++",
                              include_dirs=[numpy.get_include()],
                              define_macros=[("NPY_NO_DEPRECATED_API", "NPY_1_7_API_VERSION")],
                              extra_compile_args=["-O2", "-w"]))
setup(name="wann_ref", ext_modules=cythonize(exts, language_level=3, quiet=True,
      compiler_directives={"boundscheck": False}), script_args=["build_ext", "--inplace", "-q"])
'''


def build(force=False):
    if not os.path.isdir(REF):
        print("[build_ref] %s not present; keeping prebuilt oracle/_ref" % REF)
        return os.path.isdir(OUT)
    marker = os.path.join(OUT, ".built")
    if os.path.exists(marker) and not force:
        return True
    work = tempfile.mkdtemp(prefix="wann_ref_build_")
    try:
        for pkg in PKGS:
            shutil.copytree(os.path.join(REF, pkg), os.path.join(work, pkg))
            init = os.path.join(work, pkg, "__init__.py")
            if not os.path.exists(init):
                open(init, "w").close()
        # shim 1: drop the unused NumPy-1-only ctypedef (scratch copy only)
        up = os.path.join(work, "tools", "utils.pyx")
        src = open(up).read().replace("ctypedef np.int_t DTYPE_t", "")
        open(up, "w").write(src)
        with open(os.path.join(work, "setup_ref.py"), "w") as f:
            f.write(SETUP % {"pkgs": PKGS, "skip": SKIP})
        subprocess.check_call([sys.executable, "setup_ref.py"], cwd=work)
        if os.path.isdir(OUT):
            shutil.rmtree(OUT)
        for pkg in PKGS:
            os.makedirs(os.path.join(OUT, pkg))
            open(os.path.join(OUT, pkg, "__init__.py"), "w").close()
            for so in glob.glob(os.path.join(work, pkg, "*.so")):
                shutil.copy2(so, os.path.join(OUT, pkg))
        open(marker, "w").write("built from %s\n" % REF)
        return True
    finally:
        shutil.rmtree(work, ignore_errors=True)


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv)
    print("[build_ref] ok" if ok else "[build_ref] unavailable")
