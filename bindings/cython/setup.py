"""Builds the Cython binding in place:  python bindings/cython/setup.py build_ext --build-lib <dir>
(the reference's python/setup.py, corintho_ai/python/setup.py:17-31, with its ../cpp/src/*.cpp
sources replaced by one library)."""
import os

from Cython.Build import cythonize
from setuptools import Extension, setup

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
LIB = os.path.join(ROOT, "corintho_ai_b200")

setup(name="corintho_b200_cy",
      ext_modules=cythonize([Extension("corintho_b200_cy", [os.path.join(HERE, "corintho_b200_cy.pyx")],
                                       include_dirs=[os.path.join(ROOT, "include")], library_dirs=[LIB],
                                       libraries=["corintho_b200"], runtime_library_dirs=[LIB],
                                       extra_compile_args=["-O2", "-std=c++17"], language="c++")],
                            build_dir=os.environ.get("CB200_CY_BUILD", os.path.join(HERE, "build"))))
