"""Build the reference's only native component (the Cython ``d3rlpy.dataset``
extension) from the sources where they lie under /root/reference.

TEST INFRASTRUCTURE ONLY.  Outputs go to ``oracle/_ref/`` (git-ignored); no
reference source is copied into this repository.  Recipe follows the
reference's own ``setup.py:16-27`` (``-std=c++11 -O3 -ffast-math``, no OpenMP)
with the one shim SURVEY.md §8(c) documents: ``include_path=['d3rlpy']`` for
the pre-Cython-3 implicit-relative ``from dataset cimport CTransition``
(``d3rlpy/dataset.pyx:15``).

Usage:  python oracle/build_ref.py        (no-op when /root/reference is absent)
"""
import os
import shutil
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("D3RLPY_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")


def ext_path() -> str:
    return os.path.join(OUT, "dataset" + sysconfig.get_config_var("EXT_SUFFIX"))


def build(force: bool = False) -> str:
    pyx = os.path.join(REF, "d3rlpy", "dataset.pyx")
    if not os.path.exists(pyx):
        return ""  # GPU box: the prebuilt .so (if any) travels with the snapshot
    so = ext_path()
    if os.path.exists(so) and not force and os.path.getmtime(so) >= os.path.getmtime(pyx):
        return so
    import numpy as np
    from Cython.Build import cythonize  # noqa: F401  (presence check)

    build_dir = os.path.join(OUT, "build")
    os.makedirs(build_dir, exist_ok=True)
    cpp = os.path.join(build_dir, "dataset.cpp")
    # cythonize from the read-only tree: -o puts the generated C++ under oracle/_ref/build
    subprocess.check_call(
        [sys.executable, "-m", "cython", "--cplus", "-3", "-I", os.path.join(REF, "d3rlpy"),
         "--module-name", "d3rlpy.dataset", "-o", cpp, pyx],
        cwd=build_dir,
    )
    inc = [
        "-I" + sysconfig.get_paths()["include"],
        "-I" + np.get_include(),
        "-I" + os.path.join(REF, "d3rlpy", "cpp", "include"),
    ]
    cmd = ["g++", "-shared", "-fPIC", "-std=c++11", "-O3", "-ffast-math", "-w",
           "-DNPY_NO_DEPRECATED_API=NPY_1_7_API_VERSION"] + inc + [cpp, "-o", so]
    subprocess.check_call(cmd)
    shutil.rmtree(build_dir, ignore_errors=True)
    return so


if __name__ == "__main__":
    p = build(force="--force" in sys.argv)
    print(p if p else "reference tree not present; nothing built")
