"""ctypes loader for libgeneob200.so (built in-tree by geneo4petsc_b200/csrc/Makefile)."""
import ctypes
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgeneob200.so")
HEADER = os.path.join(os.path.dirname(_HERE), "include", "geneo_b200.h")


def header_functions():
    """Names of every function include/geneo_b200.h declares (used by the CPU-side export test)."""
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(geneo_[a-z0-9_]+)\s*\(", src)))


def load():
    if not os.path.exists(LIB_PATH):
        raise ImportError("libgeneob200.so is missing: run `make -C geneo4petsc_b200/csrc` (or __graft_entry__.build()). "
                          "There is no CPU fallback.")
    return ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)
