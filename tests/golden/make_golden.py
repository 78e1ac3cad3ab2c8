#!/usr/bin/env python
"""Generate the compact golden fixtures under tests/golden/ from the reference's own data files.

Run HERE (in the build container, where /root/reference is mounted):

    python tests/golden/make_golden.py

It reads the Amber prmtop topologies and the OpenMM State XML files that the reference's only
numerical test uses (reference python/tests/test_abfe.py:14-19,41-43,128 and
example/rbfe/rbfe.py:16,41-44,183) and writes one small ``.npz`` per system containing exactly
what the direct-space ATM hot path consumes:

    pos      (N,3) f8  nm      unwrapped State-XML positions
    box      (3,)  f8  nm      rectangular box edge lengths (State XML PeriodicBoxVectors)
    charge   (N,)  f8  e       prmtop CHARGE / 18.2223
    sigma    (N,)  f8  nm      per-type, from the diagonal LENNARD_JONES_{A,B}COEF
    epsilon  (N,)  f8  kJ/mol
    excl     (E,2) i4          excluded pairs (prmtop NUMBER_EXCLUDED_ATOMS / EXCLUDED_ATOMS_LIST)
    exc14    (X,2) i4, exc14_par (X,3) f8   1-4 exceptions: chargeProd (e^2), sigma (nm), epsilon (kJ/mol)
    displ    (N,3) f8  nm      ATM displacement of every atom as the test / example scripts set it
    lig1, lig2     i4          displaced atom indices (residue 2, residue 3)
    pin_u    f8  kJ/mol        the reference's pinned perturbation energy (test_abfe.py:148) or NaN
    params   (9,)  f8          lambda1 lambda2 alpha u0 w0 umax ubcore acore direction

and, for the second pin of the reference's test (potential energy of force groups {0, ATM}, test_abfe.py:147), the
rest of the Amber energy function as prmtop.createSystem(PME, 1 nm, constraints=HBonds) defines it
(test_abfe.py:41-43) plus the two restraints the test adds (test_abfe.py:60-83):

    pin_pe   f8  kJ/mol        -116071.0 (abfe) or NaN
    mass     (N,)  f8  amu     prmtop MASS (centroid weights of the CM-CM restraint)
    bonds    (B,2) i4, bond_par  (B,2) f8   k (kJ/mol/nm^2, E = k (r-r0)^2), r0 (nm); bonds WITHOUT hydrogen only
                                            (bonds to hydrogen are constraints, rigid water has no bond term)
    angles   (A,3) i4, angle_par (A,2) f8   k (kJ/mol/rad^2, E = k (t-t0)^2), t0 (rad)
    dihedrals (D,4) i4, dihedral_par (D,3) f8   k (kJ/mol), periodicity, phase (rad); E = k (1 + cos(n phi - phase))
    restr_ref (N,3) f8 nm      inpcrd coordinates (centres of the flat-bottom position restraints)
    posres_atoms i4            receptor carbons with index < 40 (test_abfe.py:76-83); fc 25 kcal/mol/A^2, tol 0.5 A
    cm_lig, cm_rcpt i4         atoms of the CM-CM flat-bottom restraint (kf 25 kcal/mol/A^2, tol 5 A, offset 0)

The GPU box has no /root/reference, so the tests only ever read the .npz files.
Nothing from the reference's *source code* is copied; these are data fixtures.
"""
import os
import re
import sys
import numpy as np

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def read_prmtop(path):
    """Minimal Amber7 prmtop reader: returns {flag: list of str fields} using the %FORMAT width."""
    sections = {}
    flag = None
    width = None
    with open(path) as fh:
        for line in fh:
            if line.startswith("%VERSION"):
                continue
            if line.startswith("%FLAG"):
                flag = line.split()[1]
                sections[flag] = []
                width = None
                continue
            if line.startswith("%FORMAT"):
                m = re.match(r"%FORMAT\((\d+)([aIEe])(\d+)", line)
                width = int(m.group(3))
                continue
            if line.startswith("%COMMENT"):
                continue
            line = line.rstrip("\n")
            for k in range(0, len(line), width):
                field = line[k:k + width]
                if field.strip() != "" or sections[flag] is None:
                    sections[flag].append(field)
    return sections


def ints(sec):
    return np.array([int(x) for x in sec], dtype=np.int64)


def floats(sec):
    return np.array([float(x) for x in sec], dtype=np.float64)


def read_state_xml(path):
    pos = []
    box = []
    with open(path) as fh:
        for line in fh:
            s = line.strip()
            if s.startswith("<Position "):
                m = re.findall(r'="([^"]*)"', s)
                pos.append([float(v) for v in m[:3]])
            elif s.startswith("<A ") or s.startswith("<B ") or s.startswith("<C "):
                m = re.findall(r'="([^"]*)"', s)
                box.append([float(v) for v in m[:3]])
    box = np.array(box)
    assert np.allclose(box - np.diag(np.diag(box)), 0.0), "fixtures are rectangular"
    return np.array(pos), np.diag(box).copy()


def read_inpcrd(path, natom):
    with open(path) as fh:
        lines = fh.readlines()
    assert int(lines[1].split()[0]) == natom
    vals = []
    for ln in lines[2:]:
        ln = ln.rstrip("\n")
        vals += [float(ln[k:k + 12]) for k in range(0, len(ln), 12) if ln[k:k + 12].strip()]
    return np.array(vals[:3 * natom]).reshape(natom, 3) * 0.1


def bonded_terms(top, natom):
    """The bonded part of the Amber energy function in OpenMM units (kJ/mol, nm, rad)."""
    kcal = 4.184
    out = {}
    bk, br = floats(top["BOND_FORCE_CONSTANT"]), floats(top["BOND_EQUIL_VALUE"])
    b = ints(top["BONDS_WITHOUT_HYDROGEN"]).reshape(-1, 3)
    out["bonds"] = (b[:, :2] // 3).astype(np.int32)
    out["bond_par"] = np.stack([bk[b[:, 2] - 1] * kcal * 100.0, br[b[:, 2] - 1] * 0.1], axis=1)
    ak, a0 = floats(top["ANGLE_FORCE_CONSTANT"]), floats(top["ANGLE_EQUIL_VALUE"])
    a = np.concatenate([ints(top[n]).reshape(-1, 4) for n in ("ANGLES_INC_HYDROGEN", "ANGLES_WITHOUT_HYDROGEN")])
    out["angles"] = (a[:, :3] // 3).astype(np.int32)
    out["angle_par"] = np.stack([ak[a[:, 3] - 1] * kcal, a0[a[:, 3] - 1]], axis=1)
    dk, dn, dp = (floats(top["DIHEDRAL_FORCE_CONSTANT"]), floats(top["DIHEDRAL_PERIODICITY"]),
                  floats(top["DIHEDRAL_PHASE"]))
    d = np.concatenate([ints(top[n]).reshape(-1, 5) for n in ("DIHEDRALS_INC_HYDROGEN", "DIHEDRALS_WITHOUT_HYDROGEN")])
    out["dihedrals"] = (np.abs(d[:, :4]) // 3).astype(np.int32)
    out["dihedral_par"] = np.stack([dk[d[:, 4] - 1] * kcal, dn[d[:, 4] - 1], dp[d[:, 4] - 1]], axis=1)
    out["mass"] = floats(top["MASS"])
    assert out["mass"].size == natom
    return out


def build(prmtop_path, xml_path, lig_resids, signs, params, pin_u, inpcrd_path=None, pin_pe=float("nan")):
    top = read_prmtop(prmtop_path)
    ptr = ints(top["POINTERS"])
    natom, ntypes = int(ptr[0]), int(ptr[1])
    charge = floats(top["CHARGE"]) / 18.2223
    tindex = ints(top["ATOM_TYPE_INDEX"])
    nbidx = ints(top["NONBONDED_PARM_INDEX"])
    acoef = floats(top["LENNARD_JONES_ACOEF"])
    bcoef = floats(top["LENNARD_JONES_BCOEF"])
    assert charge.size == natom and tindex.size == natom

    # per-type sigma/epsilon from the diagonal A/B (kcal/mol, Angstrom) -> nm, kJ/mol
    sig_t = np.zeros(ntypes)
    eps_t = np.zeros(ntypes)
    for t in range(ntypes):
        k = nbidx[ntypes * t + t] - 1
        a, b = acoef[k], bcoef[k]
        if a == 0.0 or b == 0.0:
            rmin, eps = 1.0, 0.0
        else:
            rmin = (2.0 * a / b) ** (1.0 / 6.0)
            eps = 0.25 * b * b / a
        sig_t[t] = rmin * 2.0 ** (-1.0 / 6.0) * 0.1
        eps_t[t] = eps * 4.184
    sigma = sig_t[tindex - 1]
    epsilon = eps_t[tindex - 1]

    # exclusions (Amber: each atom lists its higher-index partners, 1-based, a single 0 when none)
    nexc = ints(top["NUMBER_EXCLUDED_ATOMS"])
    elist = ints(top["EXCLUDED_ATOMS_LIST"])
    excl = []
    k = 0
    for i in range(natom):
        for j in elist[k:k + nexc[i]]:
            if j > 0:
                excl.append((i, int(j) - 1))
        k += nexc[i]
    excl = np.array(sorted(set(excl)), dtype=np.int32)

    # 1-4 exceptions from the dihedral lists (third index < 0 => no 1-4 term; scale factors per type)
    scee = floats(top["SCEE_SCALE_FACTOR"])
    scnb = floats(top["SCNB_SCALE_FACTOR"])
    seen = set()
    e14, p14 = [], []
    for name in ("DIHEDRALS_INC_HYDROGEN", "DIHEDRALS_WITHOUT_HYDROGEN"):
        d = ints(top[name]).reshape(-1, 5)
        for i3, j3, k3, l3, it in d:
            if k3 < 0 or l3 < 0:
                continue
            i, l = int(i3) // 3, int(l3) // 3
            key = (min(i, l), max(i, l))
            if key in seen:
                continue
            seen.add(key)
            se, sn = scee[it - 1], scnb[it - 1]
            qq = charge[i] * charge[l] / se if se != 0 else 0.0
            ee = np.sqrt(epsilon[i] * epsilon[l]) / sn if sn != 0 else 0.0
            e14.append(key)
            p14.append((qq, 0.5 * (sigma[i] + sigma[l]), ee))
    e14 = np.array(e14, dtype=np.int32).reshape(-1, 2)
    p14 = np.array(p14, dtype=np.float64).reshape(-1, 3)

    # residues -> ligand atoms
    rptr = ints(top["RESIDUE_POINTER"])
    rptr = np.append(rptr, natom + 1)
    ligs = [np.arange(rptr[r - 1] - 1, rptr[r] - 1, dtype=np.int32) for r in lig_resids]

    pos, box = read_state_xml(xml_path)
    assert pos.shape == (natom, 3)

    displ = np.zeros((natom, 3))
    d = np.array([2.2, 2.2, 2.2])  # 22 Angstrom (test_abfe.py:35, rbfe.py:32)
    for lig, s in zip(ligs, signs):
        displ[lig] = s * d
    out = dict(pos=pos, box=box, charge=charge, sigma=sigma, epsilon=epsilon, excl=excl,
               exc14=e14, exc14_par=p14, displ=displ,
               lig1=ligs[0], lig2=(ligs[1] if len(ligs) > 1 else np.zeros(0, np.int32)),
               pin_u=np.float64(pin_u), params=np.array(params, dtype=np.float64))
    out.update(bonded_terms(top, natom))
    out["pin_pe"] = np.float64(pin_pe)
    if inpcrd_path is not None:
        names = [x.strip() for x in top["ATOM_NAME"]]
        rcpt = np.arange(rptr[0] - 1, rptr[1] - 1, dtype=np.int32)   # residue 1 (test_abfe.py:33,49-52)
        out["restr_ref"] = read_inpcrd(inpcrd_path, natom)
        out["posres_atoms"] = np.array([i for i in rcpt if names[i].startswith("C") and i < 40], dtype=np.int32)
        out["cm_lig"] = ligs[0]
        out["cm_rcpt"] = rcpt
    return out


def main():
    if not os.path.isdir(REF):
        sys.exit("needs /root/reference (run in the build container)")
    kcal = 4.184
    # test_abfe.py:22-31 (lambda .5/.5, alpha 0, u0 0, w0 0, umax 200 kcal, ubcore 100 kcal, acore 1/16, dir +1)
    abfe = build(f"{REF}/python/tests/temoa-g1.prmtop", f"{REF}/python/tests/temoa-g1-equil.xml",
                 [2], [+1.0], [0.5, 0.5, 0.0, 0.0, 0.0, 200 * kcal, 100 * kcal, 0.0625, 1.0], 58.2,
                 inpcrd_path=f"{REF}/python/tests/temoa-g1.inpcrd", pin_pe=-116071.0)
    np.savez_compressed(os.path.join(OUT, "temoa_g1_abfe.npz"), **abfe)
    # rbfe.py:18-26 (umax 100 kcal, ubcore 50 kcal); no reference pin exists for this system
    rbfe = build(f"{REF}/example/rbfe/temoa-g1-g4.prmtop", f"{REF}/example/rbfe/temoa-g1-g4-equil.xml",
                 [2, 3], [+1.0, -1.0], [0.5, 0.5, 0.0, 0.0, 0.0, 100 * kcal, 50 * kcal, 0.0625, 1.0],
                 float("nan"))
    np.savez_compressed(os.path.join(OUT, "temoa_g1_g4_rbfe.npz"), **rbfe)
    for name, sysd in (("abfe", abfe), ("rbfe", rbfe)):
        print(name, "N", sysd["pos"].shape[0], "excl", sysd["excl"].shape[0], "exc14", sysd["exc14"].shape[0],
              "lig1", sysd["lig1"][[0, -1]], "lig2", sysd["lig2"][[0, -1]] if sysd["lig2"].size else None,
              "qtot %.4f" % sysd["charge"].sum(), "box", sysd["box"])


if __name__ == "__main__":
    main()
