"""Builds libintegrals_b200.so in-tree with nvcc for sm_100a (see csrc/Makefile)."""
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))


def build_library(jobs=None, verbose=False):
    """make -C csrc; returns the path of the shared library.  Cross-compiles without a GPU."""
    jobs = jobs or os.cpu_count() or 4
    lib = os.path.join(_HERE, "libintegrals_b200.so")
    objdir = os.path.normpath(os.path.join(_HERE, "..", "gpurun_out", "_build"))
    if os.path.exists(lib) and not os.path.isdir(objdir):
        # a snapshot of the repo on another machine (the built library travels, the object files do not): if nothing is newer
        # than the library there is nothing to compile -- make would rebuild all 40 translation units for want of the objects
        import glob
        srcs = glob.glob(os.path.join(_HERE, "csrc", "*.cu*")) + glob.glob(os.path.join(_HERE, "csrc", "Makefile")) + \
            glob.glob(os.path.join(_HERE, "..", "include", "*.h"))
        if srcs and max(os.path.getmtime(f) for f in srcs) <= os.path.getmtime(lib):
            return lib
    cmd = ["make", "-C", os.path.join(_HERE, "csrc"), f"-j{jobs}"]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout)
    if res.returncode != 0:
        raise RuntimeError("building libintegrals_b200.so failed (see output above)")
    return os.path.join(_HERE, "libintegrals_b200.so")
