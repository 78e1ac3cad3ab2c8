#!/usr/bin/env python
"""The evaluation part of the reference's example/abfe/abfe.py (ref: example/abfe/abfe.py:18-96,121-160) against the
Blackwell back-end: build the ATM Meta-Force for the TEMOA-G1 host-guest system (ligand = atoms 196-216 displaced by
22 Angstrom along x, y and z), set the lambda = 1/2 alchemical state, evaluate the ATM force group, print the sample
line `T lambda lambda1 lambda2 alpha u0 w0 PE u` (kcal/mol) the reference script prints, and save / reload a State
XML checkpoint.

OpenMM is not needed: the system comes from the fixture arrays extracted from the reference's Amber files
(tests/golden/temoa_g1_abfe.npz, made by tests/golden/make_golden.py), and the stand-alone Context evaluates the direct
space of the variable NonbondedForce group on the GPU (PME reciprocal space, bonded terms and the integrator stay in
OpenMM).  Needs a CUDA device -- there is no CPU fallback.

    python example/abfe/abfe_single_point.py [--cpp]      (--cpp: evaluate through the C++ ATMMetaForceImpl instead)
"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "openmm-atmmetaforce-plugin_b200", "python"))
import atmmetaforce as atm          # noqa: E402
from atmmetaforce import io         # noqa: E402

kcal = 4.184
temperature = 300.0
lmbd = 0.5
lambda1, lambda2, alpha, u0, w0coeff = lmbd, lmbd, 0.0 / kcal, 0.0 * kcal, 0.0 * kcal
umsc, ubcore, acore, direction = 200.0 * kcal, 100.0 * kcal, 0.0625, 1.0
displ = [2.2, 2.2, 2.2]                                  # 22 Angstrom, in nm
nonbonded_force_group, atmforcegroup = 1, 2

g = dict(np.load(os.path.join(ROOT, "tests", "golden", "temoa_g1_abfe.npz")))
n = g["pos"].shape[0]
lig_atoms = [int(i) for i in g["lig1"]]

atmforce = atm.ATMMetaForce(lambda1, lambda2, alpha, u0, w0coeff, umsc, ubcore, acore, direction, [nonbonded_force_group])
for i in range(n):
    atmforce.addParticle(i, 0.0, 0.0, 0.0)
for i in lig_atoms:
    atmforce.setParticleParameters(i, i, displ[0], displ[1], displ[2])
atmforce.setForceGroup(atmforcegroup)

if "--cpp" in sys.argv:
    from atmmetaforce import _atmmetaforce_core as core
    system = core.System()
    for m in g["mass"]:
        system.addParticle(float(m))
    L = g["box"]
    system.setDefaultPeriodicBoxVectors([L[0], 0, 0], [0, L[1], 0], [0, 0, L[2]])
    exc = {(int(a), int(b)): (0.0, 0.3, 0.0) for a, b in g["excl"]}          # every exception excludes its pair
    for (a, b), p in zip(g["exc14"], g["exc14_par"]):
        exc[(int(a), int(b))] = tuple(float(x) for x in p)                  # ... and 1-4 pairs carry parameters
    system.addNonbondedForce(g["charge"].tolist(), g["sigma"].tolist(), g["epsilon"].tolist(), [x for ab in exc for x in ab],
                             [x for ab in exc for x in exc[ab]], cutoff=1.0, forceGroup=nonbonded_force_group)
    force = system.addATMMetaForce(atmforce)
    context = core.Context(system)
    context.setPositions(g["pos"])
    pot_energy, forces = context.calcForcesAndEnergy(True, True, (1 << 0) | (1 << atmforcegroup))
    pert_energy = core.ATMMetaForce.getPerturbationEnergy(force, context)
    print("C++ ATMMetaForceImpl: PE(ATM group: whole NonbondedForce of state 1 + lambda2 u) = %.3f kJ/mol, u = %.4f kJ/mol "
          "(reference pin 58.2 +- 0.1), max |F| = %.1f kJ/mol/nm"
          % (pot_energy, pert_energy, np.abs(forces).max()))
    print(io.format_sample_line(temperature, lmbd, lambda1, lambda2, alpha, u0, w0coeff, pot_energy, pert_energy))
    sys.exit(0)

nonbonded = atm.NonbondedDirect(g["charge"], g["sigma"], g["epsilon"], cutoff=1.0, exclusions=g["excl"],
                                exception_pairs=g["exc14"], exception_params=g["exc14_par"], force_group=nonbonded_force_group)
context = atm.Context(atmforce, nonbonded, g["box"], precision="mixed")
context.setPositions(g["pos"])
# the reference re-applies the ATM parameters after loading a state (example/abfe/abfe.py:121-133)
for name, value in ((atmforce.Lambda1(), lambda1), (atmforce.Lambda2(), lambda2), (atmforce.Alpha(), alpha), (atmforce.U0(), u0),
                    (atmforce.W0(), w0coeff), (atmforce.Umax(), umsc), (atmforce.Ubcore(), ubcore), (atmforce.Acore(), acore),
                    (atmforce.Direction(), direction)):
    context.setParameter(name, value)

state = context.getState(getEnergy=True, getForces=True, groups={0, atmforcegroup})
pot_energy = state.getPotentialEnergy()
pert_energy = atmforce.getPerturbationEnergy(context)
print("PE of the ATM group (direct + reciprocal space + dispersion correction of state 1, + lambda2 u) = %.3f kJ/mol, "
      "perturbation energy u = %.4f kJ/mol (reference pin: 58.2 +- 0.1), max |F| = %.1f kJ/mol/nm"
      % (pot_energy, float(pert_energy), np.abs(state.getForces()).max()))

with tempfile.TemporaryDirectory() as tmp:
    out = io.SampleWriter(os.path.join(tmp, "temoa-g1.out"), temperature)
    print(out.write(context, atmforce, pot_energy, lmbd))
    out.close()
    chk = os.path.join(tmp, "temoa-g1-chk.xml")
    io.write_state_xml(chk, g["pos"], g["box"], context.getParameters())
    st = io.load_state(context, chk)
    again = context.getState(getEnergy=True, groups={0, atmforcegroup}).getPotentialEnergy()
    print("checkpoint round trip: %d positions, PE again = %.3f kJ/mol (difference %.1e)" % (len(st["positions"]), again, again - pot_energy))
context.close()
