#!/usr/bin/env python
"""The evaluation part of the reference's example/rbfe/rbfe.py (ref: example/rbfe/rbfe.py:16-38,116-133,183-212) against
the Blackwell back-end: the TEMOA host with two guests, G1 bound and G4 in the solvent; the ATM Meta-Force swaps them
(G1 atoms displaced by +22 Angstrom along x, y, z, G4 atoms by -22 Angstrom: TWO displacement groups), lambda = 1/2,
soft core umax / ubcore = 100 / 50 kcal/mol.  Prints the sample line `T lambda lambda1 lambda2 alpha u0 w0 PE u`
(kcal/mol) the reference script prints.

The system comes from the fixture arrays extracted from the reference's Amber files (tests/golden/temoa_g1_g4_rbfe.npz,
made by tests/golden/make_golden.py); the stand-alone Context evaluates the whole variable-group NonbondedForce of both
states on the GPU (direct space, PME reciprocal space, dispersion correction); bonded terms, the alignment / position
restraints (group 0) and the integrator stay in OpenMM.  Needs a CUDA device -- there is no CPU fallback.

    python example/rbfe/rbfe_single_point.py [--cpp] [--platform]
        --cpp        evaluate through the C++ ATMMetaForceImpl (fused two-state launch from host positions)
        --platform   evaluate through the compiled plugin glue on the "CUDA" platform stand-in (the reference's own
                     orchestration: two linked inner contexts + the CalcATMMetaForce kernel of libATMMetaForcePluginCUDA.so)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "openmm-atmmetaforce-plugin_b200", "python"))
import atmmetaforce as atm          # noqa: E402
from atmmetaforce import io         # noqa: E402

kcal = 4.184
temperature = 300.0
lmbd = 0.5
lambda1, lambda2, alpha, u0, w0coeff = lmbd, lmbd, 0.0 / kcal, 0.0 * kcal, 0.0 * kcal
umsc, ubcore, acore, direction = 100.0 * kcal, 50.0 * kcal, 0.0625, 1.0
displ = [2.2, 2.2, 2.2]                                  # 22 Angstrom, in nm
nonbonded_force_group, atmforcegroup = 1, 2

g = dict(np.load(os.path.join(ROOT, "tests", "golden", "temoa_g1_g4_rbfe.npz")))
n = g["pos"].shape[0]
lig1_atoms = [int(i) for i in g["lig1"]]
lig2_atoms = [int(i) for i in g["lig2"]]

atmforce = atm.ATMMetaForce(lambda1, lambda2, alpha, u0, w0coeff, umsc, ubcore, acore, direction, [nonbonded_force_group])
for i in range(n):
    atmforce.addParticle(i, 0.0, 0.0, 0.0)
for i in lig1_atoms:
    atmforce.setParticleParameters(i, i, displ[0], displ[1], displ[2])
for i in lig2_atoms:
    atmforce.setParticleParameters(i, i, -displ[0], -displ[1], -displ[2])
atmforce.setForceGroup(atmforcegroup)
print("Using ATM Meta Force plugin version = %s" % atm.ATMMETAFORCE_VERSION)

if "--cpp" in sys.argv or "--platform" in sys.argv:
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
    if "--platform" in sys.argv:
        core.registerCudaPlatform()
        core.loadPluginLibrary(os.path.join(ROOT, "openmm-atmmetaforce-plugin_b200", "libATMMetaForcePluginCUDA.so"))
        context = core.Context(system, "CUDA", {"Precision": "mixed"})
    else:
        context = core.Context(system)
    context.setPositions(g["pos"])
    pot_energy, forces = context.calcForcesAndEnergy(True, True, (1 << 0) | (1 << atmforcegroup))
    pert_energy = core.ATMMetaForce.getPerturbationEnergy(force, context)
    print("platform %s (%s): PE(ATM group) = %.3f kJ/mol, u = %.4f kJ/mol, max |F| = %.1f kJ/mol/nm"
          % (context.getPlatformName(), "kernel seam + two inner contexts" if context.usesPlatformKernel(force) else "fused two-state launch",
             pot_energy, pert_energy, np.abs(forces).max()))
else:
    nonbonded = atm.NonbondedDirect(g["charge"], g["sigma"], g["epsilon"], cutoff=1.0, exclusions=g["excl"],
                                    exception_pairs=g["exc14"], exception_params=g["exc14_par"], force_group=nonbonded_force_group)
    context = atm.Context(atmforce, nonbonded, g["box"], precision="mixed")
    context.setPositions(g["pos"])
    state = context.getState(getEnergy=True, getForces=True, groups={0, atmforcegroup})
    pot_energy = state.getPotentialEnergy()
    pert_energy = float(atmforce.getPerturbationEnergy(context))
    print("PE(ATM group) = %.3f kJ/mol, perturbation energy u = %.4f kJ/mol (oracle with the exact Ewald sum: 2.107), max |F| = %.1f kJ/mol/nm"
          % (pot_energy, pert_energy, np.abs(state.getForces()).max()))
    context.close()
print(io.format_sample_line(temperature, lmbd, lambda1, lambda2, alpha, u0, w0coeff, pot_energy, pert_energy))
