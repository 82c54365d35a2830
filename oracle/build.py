"""Builds oracle/_build/libf1o.so from oracle/c/f1o.c (gcc, OpenMP, no FMA contraction)."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "c", "f1o.c")
HDR = os.path.join(HERE, "c", "f1o.h")
OUT_DIR = os.path.join(HERE, "_build")
OUT = os.path.join(OUT_DIR, "libf1o.so")


def _gcc():
    # $CC in this image points at a gcc build without libgomp.spec; prefer the distro gcc
    for cand in ("/usr/bin/gcc", shutil.which("gcc")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("gcc not found")


def build(force=False):
    if (not force and os.path.exists(OUT)
            and os.path.getmtime(OUT) >= max(os.path.getmtime(SRC), os.path.getmtime(HDR))):
        return OUT
    os.makedirs(OUT_DIR, exist_ok=True)
    cmd = [_gcc(), "-O2", "-fPIC", "-shared", "-fopenmp", "-ffp-contract=off", "-fno-fast-math",
           "-std=c11", SRC, "-o", OUT, "-lm"]
    subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    print(build(force=True))
