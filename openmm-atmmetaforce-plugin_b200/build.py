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
SOURCES = ["atm_capi.cu", "atm_copy_merge.cu", "atm_nb.cu", "atm_hrex.cu", "atm_host.cu"]
HEADERS = [os.path.join(CSRC, f) for f in ("atm_common.cuh", "atm_nb_types.cuh", "atm_nb_lists.cuh", "atm_nb_force.cuh",
                                            "atm_nb_pme.cuh")] + [os.path.join(ROOT, "include", "atm_b200.h")]

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
    cmd = [NVCC, "--shared", "-ccbin", "/usr/bin/g++", "-o", lib] + objs + ["-lcufft", "-ldl", "-Xlinker", "-rpath=/usr/local/cuda/lib64"]
    subprocess.check_call(cmd)
    return lib


FACADE = os.path.join(HERE, "python", "atmmetaforce", "_atmmetaforce_core.so")
FACADE_SRC = [os.path.join(HERE, "openmmapi", "src", "ATMMetaForce.cpp"),
              os.path.join(HERE, "openmmapi", "src", "ATMMetaForceB200Kernel.cpp"),
              os.path.join(HERE, "openmmapi", "src", "ATMMetaForceImpl.cpp"),
              os.path.join(HERE, "serialization", "ATMMetaForceProxy.cpp"),
              os.path.join(HERE, "python", "src", "atmmetaforce_core.cpp")]


def build_facade(force=False):
    """g++: the C++ facade (ATMMetaForce class, XML proxy, kernel object) + its pybind11 binding, linked against
    libatm_b200.so, plus the stand-alone C++ serialization test."""
    import sysconfig
    import pybind11
    build(force=False)
    deps = FACADE_SRC + [os.path.join(HERE, "openmmapi", "include", f) for f in os.listdir(os.path.join(HERE, "openmmapi", "include"))]
    deps += [os.path.join(HERE, "serialization", "ATMMetaForceProxy.h"), LIB,
             os.path.join(HERE, "openmmapi", "tests", "TestATMMetaForceImpl.cpp"),
             os.path.join(HERE, "serialization", "TestSerializeATMMetaForce.cpp")]
    if not force and os.path.exists(FACADE) and all(os.path.getmtime(d) <= os.path.getmtime(FACADE) for d in deps):
        return FACADE
    inc = ["-I", os.path.join(HERE, "openmmapi", "include"), "-I", os.path.join(HERE, "serialization"),
           "-I", os.path.join(ROOT, "include"), "-I", pybind11.get_include(), "-I", sysconfig.get_paths()["include"]]
    cmd = ["/usr/bin/g++", "-std=c++17", "-O2", "-shared", "-fPIC", "-fvisibility=hidden"] + inc + FACADE_SRC + \
          ["-L", HERE, "-latm_b200", "-Wl,-rpath,$ORIGIN/../..", "-o", FACADE]
    subprocess.check_call(cmd)
    test = os.path.join(HERE, "build", "TestSerializeATMMetaForce")
    os.makedirs(os.path.dirname(test), exist_ok=True)
    cmd = ["/usr/bin/g++", "-std=c++17", "-O2"] + inc[:6] + FACADE_SRC[:4] + \
          [os.path.join(HERE, "serialization", "TestSerializeATMMetaForce.cpp"), "-L", HERE, "-latm_b200",
           "-Wl,-rpath,$ORIGIN/..", "-o", test]
    subprocess.check_call(cmd)
    test = os.path.join(HERE, "build", "TestATMMetaForceImpl")
    cmd = ["/usr/bin/g++", "-std=c++17", "-O2"] + inc[:6] + FACADE_SRC[:4] + \
          [os.path.join(HERE, "openmmapi", "tests", "TestATMMetaForceImpl.cpp"), "-L", HERE, "-latm_b200",
           "-Wl,-rpath,$ORIGIN/..", "-o", test]
    subprocess.check_call(cmd)
    return FACADE


if __name__ == "__main__":
    defs = [a[2:] for a in sys.argv[1:] if a.startswith("-D")]
    outs = [a[6:] for a in sys.argv[1:] if a.startswith("--out=")]
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, defines=defs, out=outs[0] if outs else None))
    if not outs:
        print(build_facade(force="--force" in sys.argv))
