"""Builds libtortoise_b200.so (sm_100a) in-tree with nvcc.  No GPU needed."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libtortoise_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]


def sources():
    out = []
    for f in sorted(os.listdir(CSRC)):
        if f.endswith((".cu", ".cuh", ".inc", ".h")):
            out.append(os.path.join(CSRC, f))
    out.append(os.path.join(HERE, "..", "include", "tortoise_b200.h"))
    return out


def up_to_date():
    if not os.path.exists(LIB):
        return False
    t = os.path.getmtime(LIB)
    return all(os.path.getmtime(s) <= t for s in sources())


def build(force=False, verbose=False, defines=(), out=None):
    """defines / out: a variant build (-DNAME=VALUE ...) into another file, for A/B runs through TS_B200_LIB."""
    if not force and not defines and out is None and up_to_date():
        return LIB
    out = out or LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-D" + d for d in defines] + [
        "-ccbin", "/usr/bin/g++", "-o", out, os.path.join(CSRC, "tortoise_b200.cu")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stderr)
    return out


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
