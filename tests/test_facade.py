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
