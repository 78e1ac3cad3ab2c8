"""Synthetic solvated systems and lambda schedules for the benchmark configurations.

Pure numpy/scipy host code used to make INPUTS (bench.py, tests); nothing here is on the compute path.
Units: nm, kJ/mol, e.
"""
import numpy as np

KCAL = 4.184

# TIP3P-like water
_O_Q, _H_Q = -0.834, 0.417
_O_SIG, _O_EPS = 0.315075, 0.635968
_H_SIG, _H_EPS = 0.0890899, 0.0   # sigma irrelevant (epsilon 0); Amber's loader gives rmin=1 A
_OH, _HOH = 0.09572, np.deg2rad(104.52)
_WATER_DENSITY = 33.4  # molecules / nm^3


def ewald_alpha(cutoff, tol=5e-4):
    """OpenMM's default rule alpha = sqrt(-ln(2 tol)) / r_c (tol = ewaldErrorTolerance, default 5e-4)."""
    return float(np.sqrt(-np.log(2.0 * tol)) / cutoff)


def pme_grid(box, alpha, tol=5e-4):
    """OpenMM's rule for the PME mesh: ceil(2 alpha L / (3 tol^(1/5))) rounded up to a 2/3/5/7-smooth size."""
    def legal(n):
        while True:
            m = n
            for p in (2, 3, 5, 7):
                while m % p == 0:
                    m //= p
            if m == 1:
                return n
            n += 1
    return [legal(max(6, int(np.ceil(2.0 * alpha * float(L) / (3.0 * tol ** 0.2))))) for L in box]


def _lattice_in_sphere(center, radius, n, rng, jitter=0.02):
    """n jittered simple-cubic lattice points inside a sphere (spacing chosen to fit exactly n)."""
    spacing = (4.0 / 3.0 * np.pi * radius ** 3 / n) ** (1.0 / 3.0)
    while True:
        m = int(np.ceil(radius / spacing)) + 1
        g = np.arange(-m, m + 1) * spacing
        pts = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3)
        r2 = (pts ** 2).sum(1)
        inside = pts[r2 <= radius ** 2]
        if inside.shape[0] >= n:
            order = np.argsort((inside ** 2).sum(1), kind="stable")
            pts = inside[order[:n]]
            break
        spacing *= 0.98
    pts = pts + rng.uniform(-jitter, jitter, pts.shape)
    return pts + np.asarray(center)


def _chain_exclusions(first, count):
    """1-2 and 1-3 style exclusions along the index chain of a molecule."""
    ex = []
    for i in range(count):
        for k in (1, 2):
            if i + k < count:
                ex.append((first + i, first + i + k))
    return ex


def make_system(box_edge, n_blob=0, n_lig=40, n_ligands=1, seed=12345, displ_frac=0.7, target_atoms=None):
    """Water box with an optional 'protein' blob at the centre and 1-2 ligands at its surface.

    Returns a dict with the same keys as the golden fixtures (pos, box, charge, sigma, epsilon, excl,
    exc14, exc14_par, displ, lig1, lig2).  Atom order: blob, ligand(s), waters (O,H,H).
    """
    rng = np.random.default_rng(seed)
    L = float(box_edge)
    box = np.array([L, L, L])
    centre = box / 2.0
    pos, q, sig, eps, excl = [], [], [], [], []
    n = 0
    blob_radius = 0.0
    if n_blob > 0:
        blob_radius = (n_blob / 102.0 * 3.0 / (4.0 * np.pi)) ** (1.0 / 3.0)
        bp = _lattice_in_sphere(centre, blob_radius, n_blob, rng)
        bq = rng.uniform(-0.5, 0.5, n_blob)
        bq -= bq.mean()
        pos.append(bp); q.append(bq)
        sig.append(rng.uniform(0.17, 0.23, n_blob)); eps.append(rng.uniform(0.2, 0.6, n_blob))
        excl += _chain_exclusions(0, n_blob)
        n += n_blob
    d = displ_frac * box / 2.0
    lig_idx, displ_rows = [], []
    lig_radius = (n_lig / 102.0 * 3.0 / (4.0 * np.pi)) ** (1.0 / 3.0)
    lig_centres = []
    for k in range(n_ligands):
        # ligand 0 sits at the blob surface (bound); ligand 1 sits where ligand 0 would be displaced to (bulk)
        axis = np.array([1.0, 0.0, 0.0])
        c = centre + axis * (blob_radius + lig_radius + 0.05) + (d if k == 1 else 0.0)
        lp = _lattice_in_sphere(c, lig_radius, n_lig, rng)
        lq = rng.uniform(-0.4, 0.4, n_lig)
        lq -= lq.mean()
        pos.append(lp); q.append(lq)
        sig.append(rng.uniform(0.17, 0.23, n_lig)); eps.append(rng.uniform(0.2, 0.5, n_lig))
        excl += _chain_exclusions(n, n_lig)
        lig_idx.append(np.arange(n, n + n_lig, dtype=np.int32))
        displ_rows.append((n, n_lig, d if k == 0 else -d))
        lig_centres.append(c)
        n += n_lig
    solute = np.concatenate(pos) if pos else np.zeros((0, 3))

    # waters on a jittered lattice; drop those overlapping the solute or the displaced images of the ligands
    if target_atoms is not None:
        n_w_target = max(0, (target_atoms - n) // 3)
        nl = int(np.ceil(n_w_target ** (1.0 / 3.0)))
    else:
        nl = int(round(L * _WATER_DENSITY ** (1.0 / 3.0)))
    a = L / nl
    g = (np.arange(nl) + 0.5) * a
    ow = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3)
    ow = ow + rng.uniform(-0.03, 0.03, ow.shape)
    keep = np.ones(ow.shape[0], bool)
    if solute.shape[0] > 0:
        from scipy.spatial import cKDTree
        occupied = [solute]
        for (first, cnt, dd) in displ_rows:
            occupied.append(solute[first:first + cnt] + dd)
        occ = np.concatenate(occupied)
        occ = occ - L * np.floor(occ / L)
        tree = cKDTree(occ, boxsize=L)
        dist, _ = tree.query(ow - L * np.floor(ow / L), k=1)
        keep = dist > 0.28
    ow = ow[keep]
    if target_atoms is not None:
        ow = ow[:max(0, (target_atoms - n) // 3)]
    nw = ow.shape[0]
    # random orientations
    v1 = rng.normal(size=(nw, 3)); v1 /= np.linalg.norm(v1, axis=1, keepdims=True)
    v2 = rng.normal(size=(nw, 3)); v2 -= (v2 * v1).sum(1, keepdims=True) * v1
    v2 /= np.linalg.norm(v2, axis=1, keepdims=True)
    h1 = ow + _OH * v1
    h2 = ow + _OH * (np.cos(_HOH) * v1 + np.sin(_HOH) * v2)
    wat = np.stack([ow, h1, h2], 1).reshape(-1, 3)
    pos.append(wat)
    q.append(np.tile([_O_Q, _H_Q, _H_Q], nw)); sig.append(np.tile([_O_SIG, _H_SIG, _H_SIG], nw))
    eps.append(np.tile([_O_EPS, _H_EPS, _H_EPS], nw))
    w0 = n + 3 * np.arange(nw)
    wex = np.stack([np.stack([w0, w0 + 1], 1), np.stack([w0, w0 + 2], 1), np.stack([w0 + 1, w0 + 2], 1)], 1).reshape(-1, 2)
    n += 3 * nw

    pos = np.concatenate(pos)
    charge = np.concatenate(q)
    charge[-1] -= charge.sum()  # exact neutrality (tiny correction on the last hydrogen)
    excl = np.concatenate([np.array(excl, np.int32).reshape(-1, 2), wex.astype(np.int32)])
    displ = np.zeros((n, 3))
    for (first, cnt, dd) in displ_rows:
        displ[first:first + cnt] = dd
    return dict(pos=pos, box=box, charge=charge, sigma=np.concatenate(sig), epsilon=np.concatenate(eps), excl=excl,
                exc14=np.zeros((0, 2), np.int32), exc14_par=np.zeros((0, 3)), displ=displ,
                lig1=lig_idx[0] if lig_idx else np.zeros(0, np.int32),
                lig2=lig_idx[1] if len(lig_idx) > 1 else np.zeros(0, np.int32))


def config3(seed=12345):
    """~25k-atom solvated blob + one 40-atom ligand, cutoff 0.9 nm (BASELINE.json configs[2])."""
    s = make_system(6.30, n_blob=2500, n_lig=40, n_ligands=1, seed=seed)
    s["cutoff"] = 0.9
    s["ewald_alpha"] = ewald_alpha(0.9)
    return s


def config4(seed=12345):
    """~100k-atom box, 10k-atom blob, two 40-atom ligands displaced +d / -d (BASELINE.json configs[3])."""
    s = make_system(10.0, n_blob=10000, n_lig=40, n_ligands=2, seed=seed)
    s["cutoff"] = 0.9
    s["ewald_alpha"] = ewald_alpha(0.9)
    return s


def water_box(n_atoms, n_lig=50, seed=12345):
    """Pure water + a ligand with about n_atoms atoms in total (BASELINE.json configs[4])."""
    L = (n_atoms / 3.0 / _WATER_DENSITY) ** (1.0 / 3.0)
    s = make_system(L, n_blob=0, n_lig=n_lig, n_ligands=1 if n_lig > 0 else 0, seed=seed)
    s["cutoff"] = 0.9
    s["ewald_alpha"] = ewald_alpha(0.9)
    return s


def atm_schedule_22():
    """22 lambda states of a two-leg softplus ATM schedule (the reference ships no schedule).

    Returns [22][9] rows (lambda1, lambda2, alpha, u0, w0, umax, ubcore, acore, direction).
    """
    l1 = [0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.1, 0.2, 0.3, 0.4, 0.5]
    l2 = [0.0, 0.1, 0.2, 0.3, 0.4, 0.5, 0.5, 0.5, 0.5, 0.5, 0.5]
    alpha = 0.1 / KCAL
    u0 = 110.0 * KCAL
    umax, ub, ac = 200.0 * KCAL, 100.0 * KCAL, 0.0625
    rows = []
    for a, b in zip(l1, l2):
        rows.append([a, b, alpha, u0, 0.0, umax, ub, ac, 1.0])
    for a, b in zip(l1[::-1], l2[::-1]):
        rows.append([a, b, alpha, u0, 0.0, umax, ub, ac, -1.0])
    return np.array(rows)


def partition_replicas(num_replicas, world_size):
    """Block-cyclic assignment of replicas to ranks: rank r owns replicas r, r+W, r+2W, ...."""
    return [list(range(r, num_replicas, world_size)) for r in range(world_size)]


def to_posq(sysd, padded=None, dtype=np.float32):
    """float4 posq (x,y,z,q) in atom order, zero padded to `padded` slots."""
    n = sysd["pos"].shape[0]
    P = padded or 32 * ((n + 31) // 32)
    posq = np.zeros((P, 4), dtype)
    posq[:n, :3] = sysd["pos"]
    posq[:n, 3] = sysd["charge"]
    return posq
