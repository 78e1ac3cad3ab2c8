"""numpy restatement of the parts of the Amber/OpenMM energy function that sit OUTSIDE the ATM hot path but inside the
reference's potential-energy pin (reference python/tests/test_abfe.py:147: PE of force groups {0, ATM} = -116071.0
+- 0.1 kJ/mol): harmonic bonds and angles, periodic torsions, the two flat-bottom restraints the test adds, and the
NonbondedForce's long-range dispersion correction.

TEST INFRASTRUCTURE ONLY (same rule as oracle_py.py): imported by tests/ only; the product never imports it.

The arithmetic is OpenMM's (a third-party dependency absent from /root/reference), restated from its published theory
guide ("Standard Forces": HarmonicBondForce E = k/2 (r-r0)^2 with k = 2 k_amber, HarmonicAngleForce, PeriodicTorsionForce
E = k (1 + cos(n phi - phase)), NonbondedForce long-range correction) and anchored on the reference's call sites:
  * which terms exist: prmtop.createSystem(PME, 1 nm, constraints=HBonds)  (test_abfe.py:41-43) -- bonds to hydrogen are
    constraints and carry no energy term, rigid water has none either;
  * restraints: ATMMetaForceUtils.addVsiteRestraintForceCMCM (python/ATMMetaForceUtils.py:79-141,
    "(kfcm/2)*step(d12-tolcm)*(d12-tolcm)^2", mass-weighted centroids) and addPosRestraints (:707-752,
    "0.5*fc*select(step(dist-tol), (dist-tol)^2, 0)", dist = periodicdistance to the inpcrd coordinates).
"""
import numpy as np

KCAL = 4.184


def bond_energy(pos, bonds, par):
    d = np.linalg.norm(pos[bonds[:, 0]] - pos[bonds[:, 1]], axis=1)
    return float((par[:, 0] * (d - par[:, 1]) ** 2).sum())


def angle_energy(pos, angles, par):
    v1 = pos[angles[:, 0]] - pos[angles[:, 1]]
    v2 = pos[angles[:, 2]] - pos[angles[:, 1]]
    c = (v1 * v2).sum(1) / (np.linalg.norm(v1, axis=1) * np.linalg.norm(v2, axis=1))
    th = np.arccos(np.clip(c, -1.0, 1.0))
    return float((par[:, 0] * (th - par[:, 1]) ** 2).sum())


def torsion_energy(pos, dih, par):
    b1 = pos[dih[:, 1]] - pos[dih[:, 0]]
    b2 = pos[dih[:, 2]] - pos[dih[:, 1]]
    b3 = pos[dih[:, 3]] - pos[dih[:, 2]]
    n1, n2 = np.cross(b1, b2), np.cross(b2, b3)
    m1 = np.cross(n1, b2 / np.linalg.norm(b2, axis=1)[:, None])
    phi = np.arctan2((m1 * n2).sum(1), (n1 * n2).sum(1))
    return float((par[:, 0] * (1.0 + np.cos(par[:, 1] * phi - par[:, 2]))).sum())


def cmcm_restraint_energy(pos, mass, lig, rcpt, kf, tol, offset=(0.0, 0.0, 0.0)):
    """Flat-bottom harmonic on the distance between the two mass-weighted centroids (no periodic image, as
    CustomCentroidBondForce without setUsesPeriodicBoundaryConditions)."""
    cm = lambda idx: (pos[idx] * mass[idx, None]).sum(0) / mass[idx].sum()
    d = np.linalg.norm(cm(lig) - np.asarray(offset) - cm(rcpt))
    return float(0.5 * kf * max(0.0, d - tol) ** 2)


def position_restraint_energy(pos, ref, atoms, box, fc, tol):
    dd = pos[atoms] - ref[atoms]
    dd -= box * np.round(dd / box)  # periodicdistance
    dist = np.linalg.norm(dd, axis=1)
    return float((0.5 * fc * np.where(dist > tol, (dist - tol) ** 2, 0.0)).sum())


def dispersion_correction(sigma, epsilon, cutoff, volume):
    """Isotropic long-range LJ tail: 8 pi N^2 / V * (<eps sig^12> / (9 rc^9) - <eps sig^6> / (3 rc^3)), the averages
    running over all N (N + 1) / 2 atom pairs INCLUDING i = j, Lorentz-Berthelot combination, no switching function.
    (The i = j convention moves the abfe fixture's value by 0.23 kJ/mol; the reference's PE pin discriminates: only
    this convention reproduces -116071.0 +- 0.1.)"""
    types, counts = np.unique(np.stack([sigma, epsilon], axis=1), axis=0, return_counts=True)
    n = float(sigma.size)
    sg = 0.5 * (types[:, None, 0] + types[None, :, 0])
    ep = np.sqrt(types[:, None, 1] * types[None, :, 1])
    w = np.outer(counts, counts).astype(np.float64)
    w = np.triu(w, 1) + np.diag(0.5 * counts * (counts + 1.0))
    pairs = 0.5 * n * (n + 1.0)
    s12 = (w * ep * sg ** 12).sum() / pairs
    s6 = (w * ep * sg ** 6).sum() / pairs
    return float(8.0 * n * n * np.pi * (s12 / (9.0 * cutoff ** 9) - s6 / (3.0 * cutoff ** 3)) / volume)


def group0_energy(g, pos=None):
    """Everything of the abfe fixture's force group 0: bonds + angles + torsions + both restraints (kJ/mol)."""
    pos = g["pos"] if pos is None else pos
    kf = 25.0 * KCAL * 100.0  # 25 kcal/mol/A^2 (test_abfe.py:60,71)
    parts = dict(
        bonds=bond_energy(pos, g["bonds"], g["bond_par"]),
        angles=angle_energy(pos, g["angles"], g["angle_par"]),
        torsions=torsion_energy(pos, g["dihedrals"], g["dihedral_par"]),
        cmcm=cmcm_restraint_energy(pos, g["mass"], g["cm_lig"], g["cm_rcpt"], kf, 0.5),       # tol 5 A
        posres=position_restraint_energy(pos, g["restr_ref"], g["posres_atoms"], g["box"], kf, 0.05))  # tol 0.5 A
    return sum(parts.values()), parts
