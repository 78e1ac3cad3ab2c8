#!/usr/bin/env python
"""Builds libatm_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python openmm-atmmetaforce-plugin_b200/build.py [--force] [--verbose]

The library has no Python or torch dependency; it links the static CUDA runtime.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libatm_b200.so")
SOURCES = ["atm_capi.cu", "atm_copy_merge.cu", "atm_nb.cu"]
HEADERS = [os.path.join(CSRC, "atm_common.cuh"), os.path.join(ROOT, "include", "atm_b200.h")]

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--shared", "-Xcompiler", "-fPIC",
    "-ccbin", "/usr/bin/g++",
    "-I", os.path.join(ROOT, "include"), "-I", CSRC,
]


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + HEADERS + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, defines=(), out=None):
    """defines/out: developer A/B builds (e.g. defines=["ATM_NO_NEWTON"], out="libatm_b200_exp.so")."""
    lib = os.path.join(HERE, out) if out else LIB
    if not force and not out and not needs_build():
        return LIB
    tag = ("_" + out.replace(".so", "")) if out else ""
    objs = []
    procs = []
    for s in SOURCES:
        obj = os.path.join(HERE, "build", s.replace(".cu", tag + ".o"))
        os.makedirs(os.path.dirname(obj), exist_ok=True)
        cmd = [NVCC] + [f for f in FLAGS if f != "--shared"] + ["-c", os.path.join(CSRC, s), "-o", obj] + ["-D" + x for x in defines]
        if verbose:
            cmd += ["-Xptxas", "-v"]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for cmd, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            sys.stderr.write(out)
            raise RuntimeError("nvcc failed: " + " ".join(cmd))
        if verbose:
            sys.stdout.write(out)
    cmd = [NVCC, "--shared", "-ccbin", "/usr/bin/g++", "-o", lib] + objs
    subprocess.check_call(cmd)
    return lib


if __name__ == "__main__":
    defs = [a[2:] for a in sys.argv[1:] if a.startswith("-D")]
    outs = [a[6:] for a in sys.argv[1:] if a.startswith("--out=")]
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, defines=defs, out=outs[0] if outs else None))
