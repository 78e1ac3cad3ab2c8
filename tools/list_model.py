"""CPU model of the pair-list geometry behind DESIGN.md section 3.2 ("useful-pair fraction"): for the bench system it
forms clusters the way the build kernel does (xy columns of edge (CL/rho)^(1/3), z-sorted inside a column, cut into
clusters of CL sites) and counts, per cluster, the partner SITES within cutoff + skin of at least one cluster atom (the
inner-list entries) and the (atom, partner) pairs inside the cutoff.  fill = pairs / (entries * CL) is the fraction of
the force kernel's pair slots that do useful work; it is a property of the geometry, not of the GPU.

    python tools/list_model.py [--system config3] [--sample 400]

Runs in seconds on the build container (scipy cKDTree; no GPU, no product code besides the synthetic system generator).
"""
import argparse
import os
import sys

import numpy as np
from scipy.spatial import cKDTree

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "openmm-atmmetaforce-plugin_b200", "python"))
from atmmetaforce import synthetic  # noqa: E402


def clusters_by_column(pos, box, cl):
    """(n_clusters, cl) site indices (-1 padded): xy columns, z-sorted, chunks of cl (bins padded to a multiple of cl)."""
    rho = len(pos) / np.prod(box)
    edge = (cl / rho) ** (1.0 / 3.0)
    nx, ny = max(1, int(box[0] / edge)), max(1, int(box[1] / edge))
    w = pos - box * np.floor(pos / box)
    col = np.minimum((w[:, 0] / box[0] * nx).astype(int), nx - 1) * ny + np.minimum((w[:, 1] / box[1] * ny).astype(int), ny - 1)
    order = np.lexsort((w[:, 2], col))
    out = []
    start = 0
    counts = np.bincount(col, minlength=nx * ny)
    for c in counts:
        idx = order[start:start + c]
        start += c
        for k in range(0, c, cl):
            chunk = idx[k:k + cl]
            out.append(np.pad(chunk, (0, cl - len(chunk)), constant_values=-1))
    return np.array(out)


def fill_fraction(pos, box, cl, cutoff, skin, sample, rng):
    w = pos - box * np.floor(pos / box)
    tree = cKDTree(w, boxsize=box)
    cls = clusters_by_column(pos, box, cl)
    pick = rng.choice(len(cls), size=min(sample, len(cls)), replace=False)
    entries = pairs = slots = 0
    for c in pick:
        atoms = cls[c][cls[c] >= 0]
        near = tree.query_ball_point(w[atoms], cutoff + skin)
        partners = np.unique(np.concatenate([np.asarray(x, int) for x in near]))
        partners = partners[~np.isin(partners, atoms)]
        d = w[partners][None, :, :] - w[atoms][:, None, :]
        d -= box * np.round(d / box)
        r2 = (d ** 2).sum(-1)
        entries += len(partners)
        slots += len(partners) * cl              # padding slots of a partly filled cluster are evaluated too
        pairs += int((r2 < cutoff ** 2).sum())
    return pairs / slots, entries / len(pick)


def mask_sorted_skip(pos, box, cl, cutoff, skin, sample, rng):
    """What a warp-uniform skip of cluster atoms could save: every partner carries the set of cluster atoms it is inside the
    list radius of; with the partners of a list sorted by that set and cut into 32-lane steps, a step needs only the
    atoms in the UNION of its lanes' sets.  Returns (useful fraction, executed (atom, step) fraction unsorted, sorted)."""
    w = pos - box * np.floor(pos / box)
    tree = cKDTree(w, boxsize=box)
    cls = clusters_by_column(pos, box, cl)
    pick = rng.choice(len(cls), size=min(sample, len(cls)), replace=False)
    slots = useful = plain = srt = 0
    for c in pick:
        atoms = cls[c][cls[c] >= 0]
        near = tree.query_ball_point(w[atoms], cutoff + skin)
        partners = np.unique(np.concatenate([np.asarray(x, int) for x in near]))
        partners = partners[~np.isin(partners, atoms)]
        partners = partners[rng.random(len(partners)) < 0.5]          # a pair is listed once: about half of them here
        d = w[partners][None, :, :] - w[atoms][:, None, :]
        d -= box * np.round(d / box)
        r2 = (d ** 2).sum(-1)
        inr = r2 < (cutoff + skin) ** 2
        masks = (inr * (1 << np.arange(len(atoms)))[:, None]).sum(0)
        n = len(partners)

        def executed(order):
            return sum(int(inr[:, order[k:k + 32]].any(1).sum()) * 32 for k in range(0, n, 32))
        slots += ((n + 31) // 32) * 32 * cl
        useful += int((r2 < cutoff ** 2).sum())
        plain += executed(np.arange(n))
        srt += executed(np.argsort(masks, kind="stable"))
    return useful / slots, plain / slots, srt / slots


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--system", default="config3", choices=["config3", "config4"])
    ap.add_argument("--sample", type=int, default=400, help="clusters sampled per configuration")
    args = ap.parse_args()
    s = synthetic.config3() if args.system == "config3" else synthetic.config4()
    pos, box, rc = s["pos"], np.asarray(s["box"], float), s["cutoff"]
    rng = np.random.default_rng(0)
    print(f"{args.system}: {len(pos)} atoms, box {box}, cutoff {rc} nm, density {len(pos) / np.prod(box):.1f} / nm^3")
    print("cluster size | skin (nm) | fill = pairs in cutoff / pair slots | inner-list entries per cluster")
    u, a, b = mask_sorted_skip(pos, box, 8, rc, 0.05, min(args.sample, 300), rng)
    print(f"8-atom clusters, skin 0.05: useful pair slots {u:.3f}; (atom, step) slots a warp-uniform skip would still execute: "
          f"{a:.3f} in list order, {b:.3f} with the partners sorted by their in-range atom set")
    for cl in (4, 8, 16):
        for skin in (0.0, 0.05, 0.1):
            f, e = fill_fraction(pos, box, cl, rc, skin, args.sample, rng)
            print(f"{cl:12d} | {skin:9.2f} | {f:35.3f} | {e:8.1f}")
