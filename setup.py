"""Packaging of gossipy_b200.  The sm_100a extension is built IN-TREE by ``__graft_entry__.build()`` (nvcc for the
kernels, g++ against the torch headers for the bindings / scheduler / executor):

    python setup.py build_ext --inplace        # = python -c "import __graft_entry__ as g; g.build()"   (verified)
    pip install --no-build-isolation -e .      # editable install (build_py / develop build the extension first; not exercised
                                               # in the build container, which must not write outside the repository)
"""
import os
import sys

from setuptools import Command, find_packages, setup
from setuptools.command.build_py import build_py
from setuptools.command.develop import develop

ROOT = os.path.dirname(os.path.abspath(__file__))


def _build_native() -> None:
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    import __graft_entry__
    __graft_entry__.build()


class BuildExt(Command):
    description = "compile gossipy_b200/_C*.so for sm_100a (in-tree)"
    user_options = [("inplace", "i", "ignored: the extension is always built in-tree")]
    boolean_options = ["inplace"]

    def initialize_options(self):
        self.inplace = True

    def finalize_options(self):
        pass

    def run(self):
        _build_native()


class BuildPy(build_py):
    def run(self):
        _build_native()
        super().run()


class Develop(develop):
    def run(self):
        _build_native()
        super().run()


setup(
    name="gossipy-b200",
    version="0.2.0",
    description="Blackwell-native gossip learning: gossipy's API on fused sm_100a kernels, a C++ scheduler / executor and NVLink peer memory",
    packages=find_packages(include=["gossipy_b200", "gossipy_b200.*"]),
    package_data={"gossipy_b200": ["_C*.so", "csrc/*.cpp", "csrc/*/*.cpp", "csrc/*/*.h", "csrc/*/*.cu", "csrc/*/*.cuh"]},
    python_requires=">=3.10",
    install_requires=["torch", "numpy"],
    extras_require={"data": ["scikit-learn", "scipy", "pandas"], "test": ["pytest", "hypothesis", "networkx"]},
    cmdclass={"build_ext": BuildExt, "build_py": BuildPy, "develop": Develop},
)
