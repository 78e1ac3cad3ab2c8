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
API = os.path.join(HERE, "openmmapi")
STANDIN_LIB = os.path.join(HERE, "libOpenMMStandin.so")          # plays libOpenMM.so (+ libOpenMMCUDA.so): Platform registry, Context, CUDA platform
API_LIB = os.path.join(HERE, "libATMMetaForcePlugin.so")          # ref: the API library of the same name (ATMMetaForce, Impl, proxy)
PLUGIN_LIB = os.path.join(HERE, "libATMMetaForcePluginCUDA.so")   # ref: the CUDA plugin library of the same name (kernel factory + kernel)
STANDIN_SRC = [os.path.join(API, "src", "openmm_standin_context.cpp"), os.path.join(API, "src", "openmm_standin_cuda.cpp")]
API_SRC = [os.path.join(API, "src", "ATMMetaForce.cpp"), os.path.join(API, "src", "ATMMetaForceB200Kernel.cpp"),
           os.path.join(API, "src", "ATMMetaForceImpl.cpp"), os.path.join(HERE, "serialization", "ATMMetaForceProxy.cpp")]
PLUGIN_SRC = [os.path.join(HERE, "platforms", "b200", "src", "B200ATMMetaForceKernels.cpp"),
              os.path.join(HERE, "platforms", "b200", "src", "B200ATMMetaForceKernelFactory.cpp")]
BINDING_SRC = [os.path.join(HERE, "python", "src", "atmmetaforce_core.cpp")]
TESTS = [("TestSerializeATMMetaForce", os.path.join(HERE, "serialization", "TestSerializeATMMetaForce.cpp")),
         ("TestATMMetaForceImpl", os.path.join(API, "tests", "TestATMMetaForceImpl.cpp")),
         ("TestB200ATMMetaForcePlugin", os.path.join(HERE, "platforms", "b200", "tests", "TestB200ATMMetaForcePlugin.cpp"))]
CUDA_HOME = os.environ.get("CUDA_HOME", "/usr/local/cuda")


def _headers():
    out = []
    for d in (os.path.join(API, "include"), os.path.join(HERE, "platforms", "b200", "include"), os.path.join(HERE, "serialization")):
        out += [os.path.join(d, f) for f in os.listdir(d) if f.endswith(".h")]
    return out


def build_facade(force=False):
    """g++: the host side above the C ABI, laid out like the reference's build products --
    libOpenMMStandin.so (stand-in of libOpenMM: Platform registry, Context, the CUDA-platform stand-in),
    libATMMetaForcePlugin.so (the API library: ATMMetaForce, ATMMetaForceImpl, XML proxy, kernel object),
    libATMMetaForcePluginCUDA.so (the plugin library: kernel factory + B200CalcATMMetaForceKernel, with the three
    extern "C" registration symbols), the pybind11 module and the stand-alone C++ tests."""
    import sysconfig
    import pybind11
    build(force=False)
    products = [STANDIN_LIB, API_LIB, PLUGIN_LIB, FACADE] + [os.path.join(HERE, "build", t) for t, _ in TESTS]
    deps = STANDIN_SRC + API_SRC + PLUGIN_SRC + BINDING_SRC + [src for _, src in TESTS] + _headers() + [LIB, os.path.abspath(__file__)]
    if not force and all(os.path.exists(x) for x in products) and \
            max(os.path.getmtime(d) for d in deps) <= min(os.path.getmtime(x) for x in products):
        return FACADE
    inc = ["-I", os.path.join(API, "include"), "-I", os.path.join(HERE, "platforms", "b200", "include"),
           "-I", os.path.join(HERE, "serialization"), "-I", os.path.join(ROOT, "include")]
    cxx = ["/usr/bin/g++", "-std=c++17", "-O2", "-fPIC", "-Wall", "-Wno-reorder"]
    subprocess.check_call(cxx + ["-shared"] + inc + ["-I", os.path.join(CUDA_HOME, "include")] + STANDIN_SRC +
                          ["-L", HERE, "-latm_b200", "-L", os.path.join(CUDA_HOME, "lib64"), "-lcudart", "-ldl",
                           "-Wl,-rpath,$ORIGIN", "-Wl,-rpath," + os.path.join(CUDA_HOME, "lib64"), "-o", STANDIN_LIB])
    subprocess.check_call(cxx + ["-shared"] + inc + API_SRC + ["-L", HERE, "-lOpenMMStandin", "-latm_b200", "-Wl,-rpath,$ORIGIN", "-o", API_LIB])
    subprocess.check_call(cxx + ["-shared"] + inc + PLUGIN_SRC + ["-L", HERE, "-lATMMetaForcePlugin", "-lOpenMMStandin", "-latm_b200",
                                                                  "-Wl,-rpath,$ORIGIN", "-o", PLUGIN_LIB])
    subprocess.check_call(cxx + ["-shared"] + inc + ["-I", pybind11.get_include(), "-I", sysconfig.get_paths()["include"]] + BINDING_SRC +
                          ["-L", HERE, "-lATMMetaForcePlugin", "-lOpenMMStandin", "-latm_b200", "-Wl,-rpath,$ORIGIN/../..", "-o", FACADE])
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    for name, src in TESTS:
        subprocess.check_call(cxx + inc + [src, "-L", HERE, "-lATMMetaForcePlugin", "-lOpenMMStandin", "-latm_b200", "-ldl",
                                           "-Wl,-rpath,$ORIGIN/..", "-o", os.path.join(HERE, "build", name)])
    return FACADE


if __name__ == "__main__":
    defs = [a[2:] for a in sys.argv[1:] if a.startswith("-D")]
    outs = [a[6:] for a in sys.argv[1:] if a.startswith("--out=")]
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, defines=defs, out=outs[0] if outs else None))
    if not outs:
        print(build_facade(force="--force" in sys.argv))
