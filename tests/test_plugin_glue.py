"""The compiled plugin glue without a GPU: libATMMetaForcePluginCUDA.so exports the reference's three registration
symbols (ref: platforms/cuda/src/CudaATMMetaForceKernelFactory.cpp:14-36), registers its kernel factory on the "CUDA"
platform only, registers that platform itself from the static-link entry point, and a Context on it cannot be created
without a device (no CPU fallback)."""
import ctypes
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "openmm-atmmetaforce-plugin_b200")
PLUGIN = os.path.join(PKG, "libATMMetaForcePluginCUDA.so")


def test_plugin_library_exports_the_reference_symbols():
    lib = ctypes.CDLL(PLUGIN, mode=ctypes.RTLD_GLOBAL)
    for sym in ("registerPlatforms", "registerKernelFactories", "registerATMMetaForceCudaKernelFactories"):
        assert getattr(lib, sym) is not None
    out = subprocess.run(["nm", "-D", "--defined-only", PLUGIN], capture_output=True, text=True).stdout
    exported = {ln.split()[-1] for ln in out.splitlines() if " T " in ln}
    assert {"registerPlatforms", "registerKernelFactories", "registerATMMetaForceCudaKernelFactories"} <= exported


def test_plugin_cpp_binary_registration():
    """platforms/b200/tests/TestB200ATMMetaForcePlugin.cpp in its device-free mode."""
    exe = os.path.join(PKG, "build", "TestB200ATMMetaForcePlugin")
    out = subprocess.run([exe, "cpu", PKG], capture_output=True, text=True)
    assert out.returncode == 0 and "Done" in out.stdout, out.stdout + out.stderr


def test_registration_through_the_plugin_loader():
    """Platform::loadPluginLibrary (what OpenMM does for every file in its plugin directory): without a CUDA platform the
    library registers nothing; with one, the CalcATMMetaForce kernel appears on it and only on it."""
    import torch
    from atmmetaforce import _atmmetaforce_core as core
    if "CUDA" not in core.getPlatformNames():
        core.loadPluginLibrary(PLUGIN)
        assert "CUDA" not in core.getPlatformNames()
        core.registerCudaPlatform()
    assert core.platformSupportsKernels("CUDA", ["CalcNonbondedForce"])
    core.loadPluginLibrary(PLUGIN)
    assert core.platformSupportsKernels("CUDA", ["CalcATMMetaForce"])
    assert core.getPlatformNames().count("CUDA") == 1
    with pytest.raises(Exception, match="no registered Platform"):
        core.platformSupportsKernels("HIP", ["CalcATMMetaForce"])
    if torch.cuda.is_available():
        return
    s = core.System()
    for _ in range(8):
        s.addParticle(1.0)
    s.setDefaultPeriodicBoxVectors([3, 0, 0], [0, 3, 0], [0, 0, 3])
    with pytest.raises(Exception, match="CUDA"):
        core.Context(s, "CUDA")
