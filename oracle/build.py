"""TEST INFRASTRUCTURE ONLY -- builds oracle/philox_sampler.c with gcc.

The reference is pure Python, so there is nothing to compile into
``oracle/_ref/``; the only native oracle artefact is the CPU replay of the
Philox sampler stream, built into ``oracle/_build/liboracle.so``.
"""
import ctypes
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "philox_sampler.c")
OUT_DIR = os.path.join(HERE, "_build")
OUT = os.path.join(OUT_DIR, "liboracle.so")


def build(force: bool = False) -> str:
    os.makedirs(OUT_DIR, exist_ok=True)
    if force or not os.path.exists(OUT) or os.path.getmtime(OUT) < os.path.getmtime(SRC):
        subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-std=c99", "-o", OUT, SRC])
    return OUT


_lib = None


def load() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(OUT):
            build()
        _lib = ctypes.CDLL(OUT)
    return _lib


if __name__ == "__main__":
    print(build(force=True))
