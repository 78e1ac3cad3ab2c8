"""API-surface tests of the drop-in facade (CPU): the ATMMetaForce class, its XML schema and error behaviour.
Mirrors the reference's serialization test (serialization/tests/TestSerializeATMMetaForce.cpp:12-76) value for value."""
import os
import subprocess

import numpy as np
import pytest

import atmmetaforce as atm

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _force():
    f = atm.ATMMetaForce(0.0, 0.1, 0.25, 0.5, 0.6, 200.0, 100.0, 0.07, 0.0, [1])
    f.setForceGroup(30)
    f.setName("MyATMMetaForce")
    f.addParticle(0, 0.1, 0.2, 0.3)
    f.addParticle(2, 0.4, 0.5, 0.6)
    return f


def test_serialization_round_trip_python():
    f = _force()
    xml = atm.serialize(f)
    c = atm.deserialize(xml)
    assert c.getForceGroup() == 30 and c.getName() == "MyATMMetaForce"
    assert (c.getDefaultLambda1(), c.getDefaultLambda2(), c.getDefaultAlpha(), c.getDefaultU0(), c.getDefaultW0(),
            c.getDefaultUmax(), c.getDefaultUbcore(), c.getDefaultAcore(), c.getDefaultDirection()) == \
           (0.0, 0.1, 0.25, 0.5, 0.6, 200.0, 100.0, 0.07, 0.0)
    assert c.getVariableForceGroups() == [1]
    assert c.getNumParticles() == 2
    assert c.getParticleParameters(0) == (0, 0.1, 0.2, 0.3)
    assert c.getParticleParameters(1) == (2, 0.4, 0.5, 0.6)     # the particle field is stored verbatim


def test_xml_schema_matches_reference():
    """Element / attribute / child names of serialization/src/ATMMetaForceProxy.cpp:12-38."""
    import xml.etree.ElementTree as ET
    root = ET.fromstring(atm.serialize(_force()))
    assert root.tag == "Force" and root.attrib["type"] == "ATMMetaForce" and root.attrib["version"] == "0"
    assert set(root.attrib) == {"type", "version", "forceGroup", "name", "lambda1", "lambda2", "alpha", "u0", "w0", "uMax",
                                "ubCore", "aCore", "direction"}
    assert [c.tag for c in root] == ["VariableForceGroups", "Particles"]
    assert [(p.tag, p.attrib) for p in root[0]] == [("Parameter", {"group": "1"})]
    assert [p.tag for p in root[1]] == ["Particle", "Particle"]
    assert set(root[1][0].attrib) == {"particle", "dx", "dy", "dz"}
    assert float(root.attrib["uMax"]) == 200.0 and float(root[1][1].attrib["dz"]) == 0.6


def test_unsupported_version_and_bad_index():
    with pytest.raises(atm.OpenMMException, match="Unsupported version"):
        atm.deserialize('<?xml version="1.0" ?><Force type="ATMMetaForce" version="3"/>')
    f = _force()
    with pytest.raises(atm.OpenMMException):
        f.getParticleParameters(2)
    with pytest.raises(atm.OpenMMException):
        f.setParticleParameters(-1, 0, 0, 0, 0)


def test_parameter_names_and_version():
    f = atm.ATMMetaForce
    assert [f.Lambda1(), f.Lambda2(), f.Alpha(), f.U0(), f.W0(), f.Umax(), f.Ubcore(), f.Acore(), f.Direction()] == \
        ["ATMLambda1", "ATMLambda2", "ATMAlpha", "ATMU0", "ATMW0", "ATMUmax", "ATMUbcore", "ATMAcore", "ATMDirection"]
    assert f.Version() == "0.3.1" == atm.ATMMETAFORCE_VERSION
    assert _force().usesPeriodicBoundaryConditions() is False


def test_variable_group_mask_rules():
    from atmmetaforce import _atmmetaforce_core as core
    f = _force()
    assert core.variableForceGroupsMask(f) == 2
    g = atm.ATMMetaForce(0, 0, 0, 0, 0, 1, 1, 1, 1, [3, 5])
    g.setForceGroup(5)
    with pytest.raises(atm.OpenMMException, match="cannot be one of the variable force groups"):
        core.variableForceGroupsMask(g)


def test_cpp_serialization_test_binary():
    exe = os.path.join(ROOT, "openmm-atmmetaforce-plugin_b200", "build", "TestSerializeATMMetaForce")
    if not os.path.exists(exe):
        pytest.skip("C++ test binary not built (run __graft_entry__.build())")
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and "Done" in out.stdout, out.stdout + out.stderr


# ---------------------------------------------------------------- the C++ ATMMetaForceImpl behind the OpenMM-free Context

def _cpp_system(n=64, atm_group=3, var_groups=(1,), nb_group=1, n_atm_particles=None):
    from atmmetaforce import _atmmetaforce_core as core
    s = core.System()
    for _ in range(n):
        s.addParticle(12.0)
    s.setDefaultPeriodicBoxVectors([4.0, 0, 0], [0, 4.0, 0], [0, 0, 4.0])
    s.addNonbondedForce([0.1 * (-1) ** i for i in range(n)], [0.3] * n, [0.5] * n, [0, 1, 2, 3], [0.0, 0.3, 0.0, 0.05, 0.3, 0.2],
                        cutoff=0.9, forceGroup=nb_group)
    f = atm.ATMMetaForce(0.1, 0.6, 0.2, 1.5, 0.25, 200.0, 100.0, 0.0625, 1.0, list(var_groups))
    f.setForceGroup(atm_group)
    for i in range(n if n_atm_particles is None else n_atm_particles):
        f.addParticle(i, 1.0 if i < 4 else 0.0, 0.0, 0.0)
    return core, s, s.addATMMetaForce(f)


def test_cpp_impl_registers_the_nine_parameters_and_validates():
    """ATMMetaForceImpl::initialize / getDefaultParameters (ref: openmmapi/src/ATMMetaForceImpl.cpp:69-88,130-142):
    runs on any host -- no device work happens before the first evaluation."""
    core, s, f = _cpp_system()
    ctx = core.Context(s)
    assert ctx.getParameters() == {"ATMLambda1": 0.1, "ATMLambda2": 0.6, "ATMAlpha": 0.2, "ATMU0": 1.5, "ATMW0": 0.25,
                                   "ATMUmax": 200.0, "ATMUbcore": 100.0, "ATMAcore": 0.0625, "ATMDirection": 1.0}
    ctx.setParameter("ATMLambda1", 0.3)
    assert ctx.getParameter("ATMLambda1") == 0.3
    with pytest.raises(atm.OpenMMException, match="invalid parameter name"):
        ctx.setParameter("ATMLambda3", 1.0)
    with pytest.raises(atm.OpenMMException, match="invalid parameter name"):
        ctx.getParameter("nope")
    assert core.ATMMetaForce.getPerturbationEnergy(f, ctx) == 0.0          # nothing evaluated yet
    core.ATMMetaForce.updateParametersInContext(f, ctx)                      # legal at any time
    with pytest.raises(atm.OpenMMException, match="positions have not been set"):
        ctx.calcForcesAndEnergy()
    with pytest.raises(atm.OpenMMException, match="wrong number of positions"):
        ctx.setPositions(np.zeros((3, 3)))


def test_cpp_impl_rejects_bad_systems():
    core, s, _ = _cpp_system(atm_group=1, var_groups=(1,), nb_group=2)
    with pytest.raises(atm.OpenMMException, match="cannot be one of the variable force groups"):
        core.Context(s)
    core, s, _ = _cpp_system(n_atm_particles=10)
    with pytest.raises(atm.OpenMMException, match="exactly as many particles"):
        core.Context(s)


def test_cpp_impl_skips_other_groups_and_fails_loudly_without_gpu():
    import torch
    core, s, f = _cpp_system()
    ctx = core.Context(s)
    rng = np.random.default_rng(0)
    ctx.setPositions(rng.uniform(0, 4.0, (64, 3)))
    e, frc = ctx.calcForcesAndEnergy(True, True, 1 << 1)    # the ATM group (3) is not requested: no work, no device needed
    assert e == 0.0 and frc.shape == (64, 3) and not frc.any()
    if torch.cuda.is_available():
        pytest.skip("a GPU is present (the evaluation itself is covered by tests/test_gpu_facade.py)")
    with pytest.raises(atm.OpenMMException, match="CUDA"):  # no CPU fallback: the first real evaluation needs the device
        ctx.calcForcesAndEnergy(True, True, -1)


def test_cpp_impl_test_binary():
    """openmmapi/tests/TestATMMetaForceImpl.cpp, host-logic part (parameters, names, mask rule, group skip, loud failure
    without a device)."""
    import torch
    exe = os.path.join(ROOT, "openmm-atmmetaforce-plugin_b200", "build", "TestATMMetaForceImpl")
    if not os.path.exists(exe):
        pytest.skip("C++ test binary not built (run __graft_entry__.build())")
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: tests/test_gpu_facade.py runs the binary in its gpu mode")
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and "Done" in out.stdout, out.stdout + out.stderr


def test_swig_vector_names_and_interface_file():
    """vectord / vectori (ref: python/atmmetaforceplugin.i:14-17) and the interface file itself: it declares every public
    member the reference's interface declares."""
    import os
    import re
    import atmmetaforce as atm
    v = atm.vectori([1, 2])
    v.push_back(3.0)
    assert list(v) == [1, 2, 3] and v.size() == 3 and all(type(x) is int for x in v) and not v.empty()
    d = atm.vectord(2)
    d[1] = 3
    d.append(0.5)
    assert list(d) == [0.0, 3.0, 0.5] and all(type(x) is float for x in d)
    f = atm.ATMMetaForce(0.1, 0.2, 0.0, 0.0, 0.0, 100.0, 50.0, 0.0625, 1.0, atm.vectori([1, 3]))
    g = f.getVariableForceGroups()
    assert isinstance(g, atm.vectori) and list(g) == [1, 3]
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    text = open(os.path.join(root, "openmm-atmmetaforce-plugin_b200", "python", "atmmetaforceplugin.i")).read()
    assert "%module atmmetaforce" in text and "%template(vectord) vector<double>" in text and "%template(vectori) vector<int>" in text
    header = open(os.path.join(root, "openmm-atmmetaforce-plugin_b200", "openmmapi", "include", "ATMMetaForce.h")).read()
    for name in ("getNumParticles", "addParticle", "setParticleParameters", "getParticleParameters", "updateParametersInContext",
                 "getPerturbationEnergy", "Lambda1", "Lambda2", "Alpha", "U0", "W0", "Umax", "Ubcore", "Acore", "Direction", "Version",
                 "getDefaultLambda1", "getDefaultLambda2", "getDefaultAlpha", "getDefaultU0", "getDefaultW0", "getDefaultUmax",
                 "getDefaultUbcore", "getDefaultAcore", "getDefaultDirection", "getVariableForceGroups", "cast", "isinstance"):
        assert re.search(r"\b%s\s*\(" % name, text), name
        if name not in ("cast", "isinstance"):
            assert re.search(r"\b%s\s*\(" % name, header), name


def _bond_energy_forces(pos, bonds, box=None):
    e, f = 0.0, np.zeros_like(pos)
    for a, b, r0, k in bonds:
        d = pos[b] - pos[a]
        if box is not None:
            d -= box * np.floor(d / box + 0.5)
        r = np.linalg.norm(d)
        e += 0.5 * k * (r - r0) ** 2
        f[a] += k * (r - r0) * d / r
        f[b] -= k * (r - r0) * d / r
    return e, f


def test_host_evaluated_force_standin_and_variable_group_acceptance():
    """The OpenMM-free build's HarmonicBondForce (a HostEvaluatedForce: the stand-in for OpenMM's bonded kernels) against
    numpy on the kernel-less platform, with and without the minimum-image rule, force groups honoured; and
    ATMMetaForceImpl::initialize accepts it in a variable force group next to the NonbondedForce (the generic
    variable-force hook; ref: openmmapi/src/ATMMetaForceImpl.cpp:51-65 clones every Force)."""
    from atmmetaforce import _atmmetaforce_core as core
    rng = np.random.default_rng(5)
    n = 40
    pos = rng.uniform(0, 3.0, (n, 3))
    bonds = [(int(a), int(b), float(r0), float(k)) for a, b, r0, k in
             zip(rng.integers(0, n // 2, 12), rng.integers(n // 2, n, 12), rng.uniform(0.1, 0.3, 12), rng.uniform(1e3, 3e5, 12))]
    for periodic in (False, True):
        s = core.System()
        for _ in range(n):
            s.addParticle(12.0)
        s.setDefaultPeriodicBoxVectors([3.0, 0, 0], [0, 3.0, 0], [0, 0, 3.0])
        s.addHarmonicBondForce([b[0] for b in bonds], [b[1] for b in bonds], [b[2] for b in bonds], [b[3] for b in bonds],
                               forceGroup=4, usesPeriodicBoundaryConditions=periodic)
        ctx = core.Context(s)
        ctx.setPositions(pos)
        e, f = ctx.calcForcesAndEnergy(True, True, 1 << 4)
        e_ref, f_ref = _bond_energy_forces(pos, bonds, np.full(3, 3.0) if periodic else None)
        assert abs(e - e_ref) <= 1e-12 * abs(e_ref) and np.allclose(f, f_ref, rtol=1e-12, atol=1e-9)
        e0, f0 = ctx.calcForcesAndEnergy(True, True, 1 << 3)       # another group: nothing
        assert e0 == 0.0 and not f0.any()
    with pytest.raises(atm.OpenMMException, match="Illegal particle index"):
        core.System().addHarmonicBondForce([0], [1], [0.1], [1.0])

    core, s, f = _cpp_system()          # NonbondedForce in group 1, ATM in group 3, variable groups (1,)
    s.addHarmonicBondForce([0, 1], [5, 6], [0.2, 0.2], [1000.0, 1000.0], forceGroup=1)
    ctx = core.Context(s)               # accepted: no "cannot evaluate" error
    assert not ctx.usesPlatformKernel(f)
