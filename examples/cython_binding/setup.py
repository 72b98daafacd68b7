"""python examples/cython_binding/setup.py build_ext --inplace   (needs Cython, numpy, a built libwann_b200.so)"""
import os

import numpy
from Cython.Build import cythonize
from setuptools import Extension, setup

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
LIBDIR = os.path.join(ROOT, "wann_b200")

setup(
    name="wann_eval",
    ext_modules=cythonize(
        [Extension("wann_eval", [os.path.join(HERE, "wann_eval.pyx")],
                   include_dirs=[os.path.join(ROOT, "include"), numpy.get_include()],
                   library_dirs=[LIBDIR], libraries=["wann_b200"], runtime_library_dirs=[LIBDIR],
                   define_macros=[("NPY_NO_DEPRECATED_API", "NPY_1_7_API_VERSION")])],
        include_path=[os.path.join(ROOT, "include")], language_level=3,
        build_dir=os.environ.get("WANN_CYTHON_BUILD_DIR")),          # where the generated .c goes (default: in tree)
)
