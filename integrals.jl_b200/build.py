"""Builds libintegrals_b200.so in-tree with nvcc for sm_100a (see csrc/Makefile)."""
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))


def build_library(jobs=None, verbose=False):
    """make -C csrc; returns the path of the shared library.  Cross-compiles without a GPU."""
    jobs = jobs or os.cpu_count() or 4
    cmd = ["make", "-C", os.path.join(_HERE, "csrc"), f"-j{jobs}"]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout)
    if res.returncode != 0:
        raise RuntimeError("building libintegrals_b200.so failed (see output above)")
    return os.path.join(_HERE, "libintegrals_b200.so")
