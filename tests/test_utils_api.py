"""ATMMetaForceUtils API surface (ref: python/ATMMetaForceUtils.py) against a recording stand-in for the `openmm`
module (OpenMM itself is not installable here).  Checks method names, the forces each call adds, parameter order
and unit handling; the energy expressions are additionally evaluated numerically with a tiny expression evaluator."""
import math
import sys
import types

import numpy as np
import pytest


class _Rec:
    def __init__(self, kind, n=None, expr=None):
        self.kind, self.n, self.expr = kind, n, expr
        self.params, self.groups, self.bonds, self.particles, self.torsions = [], [], [], [], []
        self.group = 0

    def addPerBondParameter(self, n): self.params.append(n)
    addPerParticleParameter = addPerTorsionParameter = addPerAngleParameter = addPerBondParameter
    def addAngle(self, a, b, c, v): self.bonds.append(((a, b, c), list(v)))
    def addGroup(self, g): self.groups.append(list(g)); return len(self.groups) - 1
    def getNumGroups(self): return len(self.groups)
    def addBond(self, p, v): self.bonds.append((list(p), list(v)))
    def addParticle(self, p, v): self.particles.append((p, list(v)))
    def addTorsion(self, a, b, c, d, v): self.torsions.append(((a, b, c, d), list(v)))
    def getNumTorsions(self): return len(self.torsions)
    def setForceGroup(self, g): self.group = g
    def getForceGroup(self): return self.group


class NonbondedForce(_Rec):
    def __init__(self, params):
        super().__init__("NonbondedForce")
        self.p = [list(x) for x in params]
    def getNumParticles(self): return len(self.p)
    def getParticleParameters(self, i): return tuple(self.p[i])
    def setParticleParameters(self, i, q, s, e): self.p[i] = [q, s, e]


class _System:
    def __init__(self, forces): self.forces = list(forces)
    def getForces(self): return self.forces
    def addForce(self, f): self.forces.append(f); return len(self.forces) - 1


@pytest.fixture()
def fake_openmm(monkeypatch):
    mod = types.ModuleType("openmm")
    for kind in ("CustomCentroidBondForce", "CustomCompoundBondForce"):
        setattr(mod, kind, (lambda k: (lambda n, expr: _Rec(k, n, expr)))(kind))
    mod.CustomTorsionForce = lambda expr: _Rec("CustomTorsionForce", 4, expr)
    mod.CustomAngleForce = lambda expr: _Rec("CustomAngleForce", 3, expr)

    class _Bond(_Rec):   # CustomBondForce.addBond(p1, p2, parameters)
        def addBond(self, p1, p2, v): self.bonds.append(((p1, p2), list(v)))
    mod.CustomBondForce = lambda expr: _Bond("CustomBondForce", 2, expr)
    mod.CustomExternalForce = lambda expr: _Rec("CustomExternalForce", 1, expr)
    monkeypatch.setitem(sys.modules, "openmm", mod)
    return mod


def _eval(expr, **vars_):
    """Evaluates an OpenMM-style expression 'value; a = ...; b = ...' (definitions after use)."""
    parts = [p.strip() for p in expr.split(";") if p.strip()]
    env = dict(vars_)
    env.update(sqrt=math.sqrt, abs=abs, floor=math.floor, atan2=math.atan2, cos=math.cos,
               step=lambda x: 1.0 if x >= 0 else 0.0, max=max)
    for p in reversed(parts[1:]):
        name, rhs = p.split("=", 1)
        env[name.strip()] = eval(rhs.replace("^", "**"), {}, env)
    return eval(parts[0].replace("^", "**"), {}, env)


def test_constructor_fixes_zero_lj_and_groups(fake_openmm):
    from atmmetaforce import ATMMetaForceUtils
    nb = NonbondedForce([(0.4, 0.0, 0.0), (-0.8, 0.3, 0.6), (0.4, 0.1, 0.0)])
    system = _System([nb])
    u = ATMMetaForceUtils(system)
    assert nb.p[0] == [0.4, 0.01, pytest.approx(4.184e-4)]     # both zero -> minimum sigma 0.1 A, epsilon 1e-4 kcal/mol
    assert nb.p[2] == [0.4, 0.1, 0.0]                           # sigma non-zero: untouched (ref: ATMMetaForceUtils.py:782)
    u.setNonbondedForceGroup(3)
    assert nb.getForceGroup() == 3


def test_cmcm_restraint(fake_openmm):
    from atmmetaforce import ATMMetaForceUtils
    system = _System([])
    u = ATMMetaForceUtils(system, fix_zero_LJparams=False)
    f = u.addVsiteRestraintForceCMCM(lig_cm_particles=[5, 6], rcpt_cm_particles=[0, 1, 2], kfcm=1000.0, tolcm=0.5, offset=[2.2, 2.2, 2.2])
    g = u.addVsiteRestraintForceCMCM([7], [0, 1], 500.0, 0.3, [0, 0, 0])
    assert f is g and f.kind == "CustomCentroidBondForce" and f.n == 2 and system.forces == [f]
    assert f.params == ["kfcm", "tolcm", "offx", "offy", "offz"]
    assert f.groups == [[5, 6], [0, 1, 2], [7], [0, 1]]
    assert f.bonds == [([0, 1], [1000.0, 0.5, 2.2, 2.2, 2.2]), ([2, 3], [500.0, 0.3, 0.0, 0.0, 0.0])]
    # flat-bottom: zero inside the tolerance, (k/2)(d - tol)^2 outside
    kw = dict(kfcm=1000.0, tolcm=0.5, offx=2.2, offy=2.2, offz=2.2, x2=0.0, y2=0.0, z2=0.0)
    assert _eval(f.expr, x1=2.5, y1=2.2, z1=2.2, **kw) == 0.0
    assert _eval(f.expr, x1=3.2, y1=2.2, z1=2.2, **kw) == pytest.approx(0.5 * 1000 * 0.5 ** 2)


def test_position_and_torsion_restraints(fake_openmm):
    from atmmetaforce import ATMMetaForceUtils
    system = _System([])
    u = ATMMetaForceUtils(system, fix_zero_LJparams=False)
    ref = np.arange(30.0).reshape(10, 3)
    f = u.addPosRestraints([1, 4], ref, fc=100.0, tol=0.05)
    assert f.kind == "CustomExternalForce" and "periodicdistance" in f.expr
    assert f.params == ["x0", "y0", "z0", "fc", "tol"]
    assert f.particles == [(1, [3.0, 4.0, 5.0, 100.0, 0.05]), (4, [12.0, 13.0, 14.0, 100.0, 0.05])]
    assert u.addPosRestraints([], ref) is None
    g = u.addPosRestraints([0], ref, periodic=False)
    assert "periodicdistance" not in g.expr
    assert _eval(g.expr, x=0.3, y=1.0, z=2.0, x0=0.0, y0=1.0, z0=2.0, fc=10.0, tol=0.1) == pytest.approx(0.5 * 10 * 0.2 ** 2)
    t = u.addTorsionalRestraintForce([0, 1, 2, 3], kphi=50.0, phi0=math.pi, phitol=0.2)
    assert t.torsions == [((0, 1, 2, 3), [50.0, math.pi, 0.2])]
    # periodic: theta = -pi + 0.1 is 0.1 away from phi0 = pi -> inside the tolerance; 0.5 away -> (k/2)(0.3)^2
    assert _eval(t.expr, theta=-math.pi + 0.1, kf=50.0, x0=math.pi, tol=0.2) == 0.0
    assert _eval(t.expr, theta=-math.pi + 0.5, kf=50.0, x0=math.pi, tol=0.2) == pytest.approx(0.5 * 50 * 0.3 ** 2)


def test_alignment_force(fake_openmm):
    from atmmetaforce import ATMMetaForceUtils
    system = _System([])
    u = ATMMetaForceUtils(system, fix_zero_LJparams=False)
    displ, theta, psi = u.addAlignmentForce([10, 11, 12], [20, 21, 22], kfdispl=250.0, ktheta=40.0, kpsi=40.0, offset=[2.2, 0, 0])
    assert system.forces == [displ, theta, psi]
    assert displ.bonds == [([20, 10], [250.0, 2.2, 0.0, 0.0])]
    assert theta.bonds == [([20, 21, 10, 11], [40.0])]
    assert psi.bonds == [([20, 21, 22, 10, 12], [20.0]), ([10, 11, 12, 20, 22], [20.0])]   # symmetrised, half each
    with pytest.raises(ValueError):
        u.addAlignmentForce([1, 2], [3, 4, 5])
    # parallel axes -> no theta penalty; antiparallel -> ktheta
    p = dict(x1=0, y1=0, z1=0, x2=0, y2=0, z2=1, x3=5, y3=0, z3=0, ktheta=40.0)
    assert _eval(theta.expr, x4=5, y4=0, z4=2, **p) == pytest.approx(0.0)
    assert _eval(theta.expr, x4=5, y4=0, z4=-2, **p) == pytest.approx(40.0)
    # roll: both reference vectors along +x about the z axis -> 0; one along -x -> kpsi (per bond: the value passed)
    q = dict(x1=0, y1=0, z1=0, x2=0, y2=0, z2=1, x3=1, y3=0, z3=0.3, x4=4, y4=4, z4=4, kpsi=20.0)
    assert _eval(psi.expr, x5=6, y5=4, z5=4.7, **q) == pytest.approx(0.0, abs=1e-12)
    assert _eval(psi.expr, x5=2, y5=4, z5=4.7, **q) == pytest.approx(20.0)


def test_cm_angle_restraints(fake_openmm):
    from atmmetaforce import ATMMetaForceUtils
    system = _System([])
    u = ATMMetaForceUtils(system, fix_zero_LJparams=False)
    lig, rcpt = [[10], [11, 12], [13]], [[0, 1], [2], [3]]
    th, ph, ps = u.addVsiteRestraintForceCMAngles(lig, rcpt, ktheta=100.0, theta0=math.radians(30), thetatol=math.radians(10),
                                                  kphi=10.0, phi0=0.5, phitol=0.2, kpsi=None)
    assert ps is None and th.kind == ph.kind == "CustomCentroidBondForce" and th.n == 4 and ph.n == 5
    assert th.groups == [[0, 1], [2], [10], [11, 12]]
    k, cos0, ctol = th.bonds[0][1]
    assert k == 100.0 and cos0 == pytest.approx(math.cos(math.radians(30)))
    # the reference's tolerance: the full span |cos(theta0 - tol) - cos(theta0 + tol)| (ref: ATMMetaForceUtils.py:612-614)
    assert ctol == pytest.approx(abs(math.cos(math.radians(20)) - math.cos(math.radians(40))))
    # the reference's group order (r1, r2, r3, l1, l2) and per-bond parameters (kf, a0 = phi0 - tol, b0 = phi0 + tol)
    assert ph.groups == [[0, 1], [2], [3], [10], [11, 12]] and ph.bonds[0][1] == pytest.approx([10.0, 0.3, 0.7])
    assert ph.params == ["kf", "a0", "b0"] and th.params == ["kf", "cos0", "ctol"]
    # phi = dihedral r3 - r2 - (r1 = l1) - l2: with r3 = (1,0,1), r2 = (0,0,1), r1 = (0,0,0) the ligand axis l2 - l1 =
    # (cos a, sin a, 0.4) has azimuth a
    ang_expr = "ang; " + ph.expr.split("ang = ", 1)[0].split("; ", 1)[1] + "ang = " + ph.expr.split("ang = ", 1)[1]
    for a in (0.3, -1.2, 2.8):
        val = _eval(ang_expr, x1=0, y1=0, z1=0, x2=0, y2=0, z2=1, x3=1, y3=0, z3=1, x4=7, y4=7, z4=7,
                    x5=7 + math.cos(a), y5=7 + math.sin(a), z5=7 - 0.4, kf=0, a0=0, b0=0)
        assert abs(val) == pytest.approx(abs(a), abs=1e-12)
    # flat bottom between a0 and b0, periodic: inside -> 0, outside -> (kf/2)(distance to the nearer edge)^2
    well = lambda ang: _eval(ph.expr.split("ang = ", 1)[0] + "ang = %r" % ang, kf=10.0, a0=0.3, b0=0.7)
    assert well(0.5) == 0.0 and well(0.31) == 0.0
    assert well(0.9) == pytest.approx(5.0 * 0.2 ** 2) and well(0.1) == pytest.approx(5.0 * 0.2 ** 2)
    assert well(0.9 + 2 * math.pi) == pytest.approx(5.0 * 0.2 ** 2)
    th2, _, ps2 = u.addVsiteRestraintForceCMAngles(lig, rcpt, ktheta=1.0, theta0=0.1, thetatol=0.1, kpsi=5.0, psi0=1.0, psitol=0.1)
    assert th2 is th and len(th.bonds) == 2 and ps2.groups == [[0, 1], [2], [10], [11, 12], [13]]
    assert ps2.bonds[0][1] == pytest.approx([5.0, 0.9, 1.1]) and ps2.params == ["kf", "a0", "b0"]


def test_boresch_restraints(fake_openmm):
    """_addVsiteRestraintForceBoresch (ref: python/ATMMetaForceUtils.py:288-384): which atoms define each of the six
    coordinates, the (kf, a0, b0) parameters, skipped terms."""
    from atmmetaforce import ATMMetaForceUtils
    system = _System([])
    u = ATMMetaForceUtils(system, fix_zero_LJparams=False)
    lig, rcpt = [10, 11, 12], [0, 1, 2]       # (A, B, C), (a, b, c)
    bond, ang, tors = u._addVsiteRestraintForceBoresch(lig, rcpt, 1000.0, 0.5, 0.1, 50.0, 1.2, 0.2, 20.0, 0.4, 0.3,
                                                       None, None, None, 21.0, -1.0, 0.25, 22.0, 2.0, 0.5)
    assert system.forces == [bond, ang, tors]
    assert bond.bonds == [((0, 10), [1000.0, 0.5, 0.1])] and bond.params == ["kf", "r0", "tol"]
    assert ang.params == tors.params == ["kf", "a0", "b0"]
    assert len(ang.bonds) == 1 and ang.bonds[0][0] == (1, 0, 10) and ang.bonds[0][1] == pytest.approx([50.0, 1.0, 1.4])   # thetaA only
    assert [t[0] for t in tors.torsions] == [(2, 1, 0, 10), (1, 0, 10, 11), (0, 10, 11, 12)]
    assert tors.torsions[1][1] == pytest.approx([21.0, -1.25, -0.75])
    # distance well: flat inside r0 +- tol, harmonic outside
    assert _eval(bond.expr, r=0.55, kf=1000.0, r0=0.5, tol=0.1) == 0.0
    assert _eval(bond.expr, r=0.8, kf=1000.0, r0=0.5, tol=0.1) == pytest.approx(500.0 * 0.2 ** 2)
    assert _eval(bond.expr, r=0.3, kf=1000.0, r0=0.5, tol=0.1) == pytest.approx(500.0 * 0.1 ** 2)
    # dihedral well is 2 pi periodic
    assert _eval(tors.expr, theta=-1.0 + 2 * math.pi, kf=21.0, a0=-1.25, b0=-0.75) == pytest.approx(0.0, abs=1e-12)
    assert _eval(tors.expr, theta=-0.5, kf=21.0, a0=-1.25, b0=-0.75) == pytest.approx(10.5 * 0.25 ** 2)
    none = u._addVsiteRestraintForceBoresch(lig, rcpt, *([None] * 18))
    assert none == (None, None, None) and len(system.forces) == 3
