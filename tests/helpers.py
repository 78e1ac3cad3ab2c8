"""Shared helpers of the parity tests (the oracle is the checker, never the thing under test)."""
import numpy as np

FIX = 4294967296.0  # 2^32 fixed point of the long force buffers

# measured parity errors of this session: {test id: {quantity: value}}; conftest.py writes them to
# gpurun_out/parity_errors.json at the end of a GPU run (the tolerances in the tests are about twice the largest
# value ever recorded there, profiles/r2_parity_errors.md)
PARITY_LOG = {}


def record(**errors):
    import os
    test = os.environ.get("PYTEST_CURRENT_TEST", "?").split(" ")[0]
    PARITY_LOG.setdefault(test, []).append({k: float(v) for k, v in errors.items()})


def oracle_system(O, s, cutoff=None, alpha=None):
    cutoff = cutoff if cutoff is not None else s.get("cutoff", 1.0)
    alpha = alpha if alpha is not None else s.get("ewald_alpha", O.ewald_alpha(cutoff))
    return O.System(s["charge"], s["sigma"], s["epsilon"], s["box"], cutoff, alpha, s["excl"], s["exc14"], s["exc14_par"])


def rel_rms(a, b):
    """relative RMS deviation of per-atom force vectors: sqrt(<|a-b|^2> / <|b|^2>)"""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.sqrt(((a - b) ** 2).sum() / (b ** 2).sum()))


def force_from_fixed(buf, n, padded):
    """[3P] int64 SoA -> (n,3) float64"""
    f = np.asarray(buf, np.int64).reshape(3, padded)[:, :n].T
    return f.astype(np.float64) / FIX


def make_backend(atm, s, cutoff, alpha, params, skin=0.1, replicas=1, perm=None):
    """Backend + posq tensor for a system dict; perm = slot -> atom permutation (OpenMM's atom reordering)."""
    import torch
    n = s["pos"].shape[0]
    be = atm.ATMBackend(n, precision="mixed", num_replicas=replicas)
    P = be.P
    atom_index = np.arange(n, dtype=np.int32) if perm is None else np.asarray(perm, np.int32)
    be.set_displacements(s["displ"], atom_index=atom_index)
    be.set_box(s["box"])
    be.set_parameters(params)
    be.nb_setup(s["charge"], s["sigma"], s["epsilon"], cutoff, alpha, skin=skin, exclusions=s["excl"],
                exception_pairs=s["exc14"], exception_params=s["exc14_par"])
    posq = np.zeros((replicas, P, 4), np.float32)
    posq[:, :n, :3] = s["pos"][atom_index]
    posq[:, :n, 3] = s["charge"][atom_index]
    return be, torch.from_numpy(posq).cuda(), atom_index
