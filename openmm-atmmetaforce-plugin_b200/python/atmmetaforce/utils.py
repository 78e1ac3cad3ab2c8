"""ATMMetaForceUtils -- helper that decorates an OpenMM System with the restraints ATM calculations use.

Same class name, method names, keyword arguments and return values as the reference helper
(ref: python/ATMMetaForceUtils.py:8-787), re-implemented on OpenMM's built-in `distance` / `angle` / `dihedral`
functions of the Custom*Forces.  Every force created here is an ordinary OpenMM force in force group 0: none of it is
evaluated by the ATM kernels.  It needs the `openmm` package at call time; OpenMM is not
installable in the build image, so only the bookkeeping is unit-tested (tests/test_utils_api.py, against a recording
stand-in for the `openmm` module).

Units: arguments may be OpenMM Quantities or plain numbers already in the MD unit system (nm, kJ/mol, radians).
"""
import math
import re


def _mm():
    try:
        import openmm
        return openmm
    except ImportError:  # older installs
        from simtk import openmm
        return openmm


def _val(x, unit_name=None):
    """Strip a Quantity to the MD unit system; pass plain numbers through."""
    if x is None:
        return None
    if hasattr(x, "value_in_unit_system"):
        try:
            from openmm import unit
        except ImportError:
            from simtk import unit
        return x.value_in_unit_system(unit.md_unit_system)
    return x


def _vec3(v):
    v = _val(v)
    return [float(_val(v[0])), float(_val(v[1])), float(_val(v[2]))]


# flat-bottom harmonic wells -----------------------------------------------------------------------------------------
# distance-like variable d >= 0:        (k/2) max(0, d - tol)^2
_FB_DIST = "0.5*kf*step(dd)*dd^2; dd = d - tol; "
# periodic angle variable (period 2 pi): (k/2) max(0, |wrap(x - x0)| - tol)^2
_FB_ANGLE = ("0.5*kf*step(da)*da^2; da = abs(dw) - tol; dw = dx - twopi*floor(dx/twopi + 0.5); dx = ang - x0; "
             "twopi = %.17g; " % (2.0 * math.pi))
# the same well written with its two edges a0 = x0 - tol, b0 = x0 + tol (the reference's per-bond parameters of the phi / psi
# restraints, ref: python/ATMMetaForceUtils.py:638-640, 684-686: code that edits the bonds afterwards relies on them)
_FB_ANGLE_AB = ("0.5*kf*step(da)*da^2; da = abs(dw) - 0.5*(b0 - a0); dw = dx - twopi*floor(dx/twopi + 0.5); dx = ang - 0.5*(a0 + b0); "
                "twopi = %.17g; " % (2.0 * math.pi))
# cosine variable:                        (k/2) max(0, |c - c0| - ctol)^2
_FB_COS = "0.5*kf*step(dc)*dc^2; dc = abs(cost - cos0) - ctol; "


class ATMMetaForceUtils(object):
    """ATM Meta Force python utilities."""

    def __init__(self, system, fix_zero_LJparams=True):
        self.system = system
        self.CMCMDistForce = None
        self.CMAngleThetaForce = None
        self.CMAnglePhiForce = None
        self.CMAnglePsiForce = None
        self.TorsionalRestraintForce = None
        if fix_zero_LJparams:
            for force in self._nonbonded_forces():
                self.fixZeroLJParams(force)

    # -- non-bonded bookkeeping ------------------------------------------------------------------------------------
    def _nonbonded_forces(self):
        pat = re.compile(".*Nonbonded.*")
        return [f for f in self.system.getForces() if pat.match(str(type(f)))]

    def setNonbondedForceGroup(self, group):
        """Places every non-bonded Force of the System in `group` (it becomes an ATM variable force group)."""
        for force in self._nonbonded_forces():
            force.setForceGroup(group)

    def fixZeroLJParams(self, force, minsigma=0.01, minepsilon=1.0e-4 * 4.184):
        """Atoms whose sigma AND epsilon are both (numerically) zero get minimum LJ parameters, so that a bare charge
        cannot sit on top of another charge when the ligand is displaced.  Defaults: 0.1 Angstrom, 1e-4 kcal/mol."""
        small = 1.0e-6
        if not hasattr(force, "getParticleParameters") or not hasattr(force, "getNumParticles"):
            return
        if "Nonbonded" not in str(type(force)) or "Custom" in str(type(force)):
            return
        for i in range(force.getNumParticles()):
            charge, sigma, epsilon = force.getParticleParameters(i)[:3]
            if _val(sigma) < small and _val(epsilon) < small:
                force.setParticleParameters(i, charge, _val(minsigma), _val(minepsilon))

    # -- centre-of-mass distance restraint (defines the binding site) -----------------------------------------------
    def addRestraintForce(self, lig_cm_particles=None, rcpt_cm_particles=None, kfcm=0.0, tolcm=0.0, offset=(0., 0., 0.)):
        """Deprecated alias of addVsiteRestraintForceCMCM()."""
        print("warning: AddRestraintForce() is deprecated. Use addVsiteRestraintForceCMCM()")
        return self.addVsiteRestraintForceCMCM(lig_cm_particles, rcpt_cm_particles, kfcm, tolcm, offset)

    def addVsiteRestraintForceCMCM(self, lig_cm_particles=None, rcpt_cm_particles=None, kfcm=0.0, tolcm=0.0,
                                   offset=(0., 0., 0.)):
        """Flat-bottom harmonic restraint on |CM(ligand) - offset - CM(receptor)|.  Returns the CustomCentroidBondForce
        (one force object is shared by all CM-CM restraints of the System)."""
        assert lig_cm_particles is not None and len(lig_cm_particles) > 0
        assert rcpt_cm_particles is not None and len(rcpt_cm_particles) > 0
        mm = _mm()
        if self.CMCMDistForce is None:
            expr = _FB_DIST.replace("tol", "tolcm").replace("kf", "kfcm") + \
                "d = sqrt((x1-offx-x2)^2 + (y1-offy-y2)^2 + (z1-offz-z2)^2)"
            force = mm.CustomCentroidBondForce(2, expr)
            for name in ("kfcm", "tolcm", "offx", "offy", "offz"):
                force.addPerBondParameter(name)
            self.system.addForce(force)
            self.CMCMDistForce = force
        force = self.CMCMDistForce
        first = force.getNumGroups()
        force.addGroup(list(lig_cm_particles))
        force.addGroup(list(rcpt_cm_particles))
        force.addBond([first, first + 1], [float(_val(kfcm)), float(_val(tolcm))] + _vec3(offset))
        return force

    # -- torsional restraint ------------------------------------------------------------------------------------------
    def addTorsionalRestraintForce(self, particles, kphi, phi0, phitol):
        """Flat-bottom harmonic restraint (2 pi periodic) on the dihedral of four particles."""
        mm = _mm()
        if self.TorsionalRestraintForce is None:
            force = mm.CustomTorsionForce(_FB_ANGLE.replace("ang", "theta"))
            for name in ("kf", "x0", "tol"):
                force.addPerTorsionParameter(name)
            self.system.addForce(force)
            self.TorsionalRestraintForce = force
        force = self.TorsionalRestraintForce
        force.addTorsion(particles[0], particles[1], particles[2], particles[3],
                         [float(_val(kphi)), float(_val(phi0)), float(_val(phitol))])
        return force

    # -- ligand-ligand alignment (RBFE) ---------------------------------------------------------------------------------
    def addAlignmentForce(self, liga_ref_particles=None, ligb_ref_particles=None, kfdispl=0.0, ktheta=0.0, kpsi=0.0,
                          offset=(0., 0., 0.)):
        """Keeps two ligands superimposed (after removing `offset`) through three reference atoms each, (a1,a2,a3) and
        (b1,b2,b3): a harmonic term on b1 - offset - a1, (ktheta/2)(1 - cos theta) between the axes a2-a1 and b2-b1, and
        (kpsi/2)(1 - cos psi) on the roll about that axis, symmetrised in a <-> b.  Returns the three forces."""
        if liga_ref_particles is None or ligb_ref_particles is None or \
                len(liga_ref_particles) != 3 or len(ligb_ref_particles) != 3:
            raise ValueError("Invalid lists of reference atoms")
        mm = _mm()
        a1, a2, a3 = liga_ref_particles
        b1, b2, b3 = ligb_ref_particles
        displ = mm.CustomCompoundBondForce(2, "0.5*kfdispl*((x1-offx-x2)^2 + (y1-offy-y2)^2 + (z1-offz-z2)^2)")
        for name in ("kfdispl", "offx", "offy", "offz"):
            displ.addPerBondParameter(name)
        displ.addBond([b1, a1], [float(_val(kfdispl))] + _vec3(offset))
        self.system.addForce(displ)

        # Lepton resolves a name from the definitions that FOLLOW its use, so un precedes ux, uy, uz
        axis = ("un = sqrt(ux^2+uy^2+uz^2); ux = x2-x1; uy = y2-y1; uz = z2-z1; ")
        theta = mm.CustomCompoundBondForce(
            4, "0.5*ktheta*(1 - (ux*vx+uy*vy+uz*vz)/(un*vn)); vn = sqrt(vx^2+vy^2+vz^2); "
               "vx = x4-x3; vy = y4-y3; vz = z4-z3; " + axis.rstrip("; "))
        theta.addPerBondParameter("ktheta")
        theta.addBond([b1, b2, a1, a2], [float(_val(ktheta))])
        self.system.addForce(theta)

        # roll: angle between the components of (p3 - p1) and (p5 - p4) perpendicular to the axis p2 - p1
        psi_expr = ("0.5*kpsi*(1 - (px*qx+py*qy+pz*qz)/(pn*qn)); pn = sqrt(px^2+py^2+pz^2); qn = sqrt(qx^2+qy^2+qz^2); "
                    "px = ax-pa*ux/un; py = ay-pa*uy/un; pz = az-pa*uz/un; pa = (ax*ux+ay*uy+az*uz)/un; "
                    "qx = bx-qb*ux/un; qy = by-qb*uy/un; qz = bz-qb*uz/un; qb = (bx*ux+by*uy+bz*uz)/un; "
                    "ax = x3-x1; ay = y3-y1; az = z3-z1; bx = x5-x4; by = y5-y4; bz = z5-z4; " + axis.rstrip("; "))
        psi = mm.CustomCompoundBondForce(5, psi_expr)
        psi.addPerBondParameter("kpsi")
        half = 0.5 * float(_val(kpsi))
        psi.addBond([b1, b2, b3, a1, a3], [half])
        psi.addBond([a1, a2, a3, b1, b3], [half])
        self.system.addForce(psi)
        return (displ, theta, psi)

    # -- ligand orientation with respect to the receptor ----------------------------------------------------------------
    def addVsiteRestraintForceCMAngles(self, lig_cm_groups=None, rcpt_cm_groups=None, ktheta=None, theta0=None,
                                       thetatol=None, kphi=None, phi0=None, phitol=None, kpsi=None, psi0=None, psitol=None):
        """Flat-bottom restraints on the orientation of the ligand frame (centroids l1,l2,l3) in the receptor frame
        (centroids r1,r2,r3): theta = angle between r2-r1 and l2-l1 (restrained in cos theta), phi = dihedral
        r3-r2-(r1=l1)-l2 (azimuth of the ligand axis), psi = dihedral r2-(r1=l1)-l2-l3 (twist about the ligand axis).
        Returns (thetaforce, phiforce, psiforce); a term whose force constant is None is skipped (None is returned)."""
        assert lig_cm_groups is not None and len(lig_cm_groups) == 3
        assert rcpt_cm_groups is not None and len(rcpt_cm_groups) == 3
        mm = _mm()
        l1, l2, l3 = (list(g) for g in lig_cm_groups)
        r1, r2, r3 = (list(g) for g in rcpt_cm_groups)
        thetaforce = phiforce = psiforce = None

        if ktheta is not None:
            if self.CMAngleThetaForce is None:
                expr = _FB_COS + ("cost = (ux*vx+uy*vy+uz*vz)/(sqrt(ux^2+uy^2+uz^2)*sqrt(vx^2+vy^2+vz^2)); "
                                  "ux = x2-x1; uy = y2-y1; uz = z2-z1; vx = x4-x3; vy = y4-y3; vz = z4-z3")
                f = mm.CustomCentroidBondForce(4, expr)
                for name in ("kf", "cos0", "ctol"):
                    f.addPerBondParameter(name)
                self.system.addForce(f)
                self.CMAngleThetaForce = f
            thetaforce = self.CMAngleThetaForce
            g0 = thetaforce.getNumGroups()
            for g in (r1, r2, l1, l2):
                thetaforce.addGroup(g)
            t0, tt = float(_val(theta0)), float(_val(thetatol))
            cos0 = math.cos(t0)
            # the reference's tolerance is the full span of cos(theta) over [theta0 - tol, theta0 + tol] clipped to [0, pi]
            # (ref: python/ATMMetaForceUtils.py:612-614)
            a0, b0 = max(0.0, t0 - tt), min(math.pi, t0 + tt)
            ctol = abs(math.cos(a0) - math.cos(b0))
            thetaforce.addBond([g0, g0 + 1, g0 + 2, g0 + 3], [float(_val(ktheta)), cos0, ctol])

        def dihedral_force(attr, groups, points, k, x0, tol):
            # Group order and per-bond parameters (kf, a0 = x0 - tol, b0 = x0 + tol) are the reference's
            # (ref: :642-656 phi: r1, r2, r3, l1, l2; :688-702 psi: r1, r2, l1, l2, l3).  `points` names, in terms of the
            # five centroids p1..p5, the three bond vectors b1, b2, b3 of the dihedral once the two frames share an origin.
            f = getattr(self, attr)
            if f is None:
                # (upstream's phi expression tests the variable `psi` instead of `phi`, ref: :638-641 -- a typo that makes its
                # phi restraint fail to compile; the intended dihedral is implemented here)
                expr = _FB_ANGLE_AB + (
                    "ang = atan2(sy, cx); "
                    # IUPAC dihedral atan2(|b2| b1.(b2 x b3), (b1 x b2).(b2 x b3))
                    "sy = bn*(b1x*n2x + b1y*n2y + b1z*n2z); cx = n1x*n2x + n1y*n2y + n1z*n2z; "
                    "n1x = b1y*b2z-b1z*b2y; n1y = b1z*b2x-b1x*b2z; n1z = b1x*b2y-b1y*b2x; "
                    "n2x = b2y*b3z-b2z*b3y; n2y = b2z*b3x-b2x*b3z; n2z = b2x*b3y-b2y*b3x; "
                    "bn = sqrt(b2x^2+b2y^2+b2z^2); " + points)
                f = mm.CustomCentroidBondForce(5, expr)
                for name in ("kf", "a0", "b0"):
                    f.addPerBondParameter(name)
                self.system.addForce(f)
                setattr(self, attr, f)
            g0 = f.getNumGroups()
            for g in groups:
                f.addGroup(g)
            c, t = float(_val(x0)), float(_val(tol))
            f.addBond([g0 + i for i in range(5)], [float(_val(k)), c - t, c + t])
            return f

        def vec(name, a, b):   # name = p_a - p_b, component-wise
            return "; ".join(f"{name}{c} = {c}{a}-{c}{b}" for c in "xyz")

        if kphi is not None:
            # groups (r1, r2, r3, l1, l2); dihedral r3 - r2 - (r1 = l1) - l2: b1 = r2 - r3, b2 = r1 - r2, b3 = l2 - l1
            phiforce = dihedral_force("CMAnglePhiForce", (r1, r2, r3, l1, l2),
                                      "; ".join((vec("b1", 2, 3), vec("b2", 1, 2), vec("b3", 5, 4))), kphi, phi0, phitol)
        if kpsi is not None:
            # groups (r1, r2, l1, l2, l3); dihedral r2 - (r1 = l1) - l2 - l3: b1 = r1 - r2, b2 = l2 - l1, b3 = l3 - l2
            psiforce = dihedral_force("CMAnglePsiForce", (r1, r2, l1, l2, l3),
                                      "; ".join((vec("b1", 1, 2), vec("b2", 4, 3), vec("b3", 5, 4))), kpsi, psi0, psitol)
        return (thetaforce, phiforce, psiforce)

    # -- Boresch-style restraints between three receptor atoms (a, b, c) and three ligand atoms (A, B, C) ---------------------
    def _addVsiteRestraintForceBoresch(self, lig_ref_particles, rcpt_ref_particles, kfrA, rA0, rAtol, kfthA, thA0, thAtol,
                                       kfphA, phA0, phAtol, kfthB, thB0, thBtol, kfphB, phB0, phBtol, kfphC, phC0, phCtol):
        """Flat-bottom wells on the six Boresch coordinates (ref: python/ATMMetaForceUtils.py:280-384; private and unused
        upstream, kept for API parity): rA = |a A|, thetaA = angle(b, a, A), thetaB = angle(a, A, B),
        phiA = dihedral(c, b, a, A), phiB = dihedral(b, a, A, B), phiC = dihedral(a, A, B, C).  A term whose force
        constant is None is skipped.  Returns (CustomBondForce, CustomAngleForce, CustomTorsionForce), None where unused.
        Per-term parameters as upstream: (kf, r0, tol) for the distance, (kf, a0, b0) = centre -+ tolerance for the rest.
        Angles use period pi in the wrap, dihedrals 2 pi, as upstream."""
        assert len(lig_ref_particles) == 3 and len(rcpt_ref_particles) == 3
        mm = _mm()
        A, B, C = lig_ref_particles
        a, b, c = rcpt_ref_particles
        bondforce = angleforce = torsforce = None
        if kfrA is not None:
            bondforce = mm.CustomBondForce("0.5*kf*step(dd)*dd^2; dd = abs(r - r0) - tol")
            for name in ("kf", "r0", "tol"):
                bondforce.addPerBondParameter(name)
            bondforce.addBond(a, A, [float(_val(kfrA)), float(_val(rA0)), float(_val(rAtol))])
            self.system.addForce(bondforce)

        def edges(x0, tol):
            x0, tol = float(_val(x0)), float(_val(tol))
            return x0 - tol, x0 + tol

        angles = [(kfthA, thA0, thAtol, (b, a, A)), (kfthB, thB0, thBtol, (a, A, B))]
        if any(k is not None for k, *_ in angles):
            angleforce = mm.CustomAngleForce(_FB_ANGLE_AB.replace("ang", "theta").replace("twopi = %.17g" % (2.0 * math.pi),
                                                                                            "twopi = %.17g" % math.pi))
            for name in ("kf", "a0", "b0"):
                angleforce.addPerAngleParameter(name)
            for k, x0, tol, (p0, p1, p2) in angles:
                if k is not None:
                    angleforce.addAngle(p0, p1, p2, [float(_val(k)), *edges(x0, tol)])
            self.system.addForce(angleforce)
        torsions = [(kfphA, phA0, phAtol, (c, b, a, A)), (kfphB, phB0, phBtol, (b, a, A, B)), (kfphC, phC0, phCtol, (a, A, B, C))]
        if any(k is not None for k, *_ in torsions):
            torsforce = mm.CustomTorsionForce(_FB_ANGLE_AB.replace("ang", "theta"))
            for name in ("kf", "a0", "b0"):
                torsforce.addPerTorsionParameter(name)
            for k, x0, tol, (p0, p1, p2, p3) in torsions:
                if k is not None:
                    torsforce.addTorsion(p0, p1, p2, p3, [float(_val(k)), *edges(x0, tol)])
            self.system.addForce(torsforce)
        return (bondforce, angleforce, torsforce)

    # -- positional restraints -------------------------------------------------------------------------------------------
    def addPosRestraints(self, particles, refpos, fc=25.0 * 4.184 * 100.0, tol=0.05, periodic=True):
        """Flat-bottom harmonic position restraints of `particles` to refpos[p] (refpos holds ALL atoms of the System).
        Defaults: 25 kcal/mol/A^2 and 0.5 A.  Returns the CustomExternalForce (None for an empty selection)."""
        if not particles or len(particles) == 0:
            return None
        mm = _mm()
        dist = "periodicdistance(x,y,z,x0,y0,z0)" if periodic else "sqrt((x-x0)^2+(y-y0)^2+(z-z0)^2)"
        force = mm.CustomExternalForce(_FB_DIST.replace("kf", "fc") + "d = " + dist)
        for name in ("x0", "y0", "z0", "fc", "tol"):
            force.addPerParticleParameter(name)
        self.system.addForce(force)
        k, t = float(_val(fc)), float(_val(tol))
        for p in particles:
            force.addParticle(p, _vec3(refpos[p]) + [k, t])
        return force
