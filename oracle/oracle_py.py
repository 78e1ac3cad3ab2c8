"""ctypes loader for the CPU oracle (oracle/libatm_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the cpu_baseline /
--impl reference legs of bench.py.  Nothing under openmm-atmmetaforce-plugin_b200/ imports this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libatm_oracle.so")


def build(force=False):
    src = os.path.join(_HERE, "atm_oracle.c")
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE], stdout=subprocess.DEVNULL)
    return _LIB


class _Sys(C.Structure):
    _fields_ = [("n", C.c_int), ("charge", C.c_void_p), ("sigma", C.c_void_p), ("epsilon", C.c_void_p),
                ("n_excl", C.c_int), ("excl", C.c_void_p), ("n_exc14", C.c_int), ("exc14", C.c_void_p),
                ("exc14_par", C.c_void_p), ("box", C.c_double * 3), ("cutoff", C.c_double),
                ("ewald_alpha", C.c_double)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB)
        _lib.atm_oracle_softcore.restype = C.c_double
        _lib.atm_oracle_softcore.argtypes = [C.c_double] * 4 + [C.POINTER(C.c_double)]
        _lib.atm_oracle_nb_direct.restype = C.c_double
        _lib.atm_oracle_ewald_recip.restype = C.c_double
        _lib.atm_oracle_ewald_recip.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_void_p]
        _lib.atm_oracle_pme_recip.restype = C.c_double
        _lib.atm_oracle_pme_recip.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        _lib.atm_oracle_num_threads.restype = C.c_int
        _lib.atm_oracle_merge_ref.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_double]
        _lib.atm_oracle_hybrid_force_i64.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double]
        _lib.atm_oracle_step.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p,
                                         C.c_void_p]
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def ewald_alpha(cutoff, tol=5e-4):
    """OpenMM's rule for the Ewald splitting parameter: alpha = sqrt(-ln(2 tol)) / r_c."""
    return float(np.sqrt(-np.log(2.0 * tol)) / cutoff)


class System:
    """Keeps the numpy arrays alive next to the C struct."""

    def __init__(self, charge, sigma, epsilon, box, cutoff, alpha, excl=None, exc14=None, exc14_par=None):
        self.charge = np.ascontiguousarray(charge, np.float64)
        self.sigma = np.ascontiguousarray(sigma, np.float64)
        self.epsilon = np.ascontiguousarray(epsilon, np.float64)
        self.excl = np.ascontiguousarray(excl if excl is not None else np.zeros((0, 2)), np.int32)
        self.exc14 = np.ascontiguousarray(exc14 if exc14 is not None else np.zeros((0, 2)), np.int32)
        self.exc14_par = np.ascontiguousarray(exc14_par if exc14_par is not None else np.zeros((0, 3)), np.float64)
        self.n = self.charge.size
        s = _Sys()
        s.n = self.n
        s.charge, s.sigma, s.epsilon = _p(self.charge), _p(self.sigma), _p(self.epsilon)
        s.n_excl, s.excl = self.excl.shape[0], _p(self.excl)
        s.n_exc14, s.exc14, s.exc14_par = self.exc14.shape[0], _p(self.exc14), _p(self.exc14_par)
        s.box[0], s.box[1], s.box[2] = [float(b) for b in box]
        s.cutoff, s.ewald_alpha = float(cutoff), float(alpha)
        self.c = s

    def nb_direct(self, pos, want_force=True):
        pos = np.ascontiguousarray(pos, np.float64)
        f = np.zeros_like(pos) if want_force else None
        out = np.zeros(4)
        e = lib().atm_oracle_nb_direct(C.byref(self.c), _p(pos), _p(f), _p(out))
        return e, out, f

    def ewald_recip(self, pos, tol=1e-12, want_force=False):
        pos = np.ascontiguousarray(pos, np.float64)
        f = np.zeros_like(pos) if want_force else None
        e = lib().atm_oracle_ewald_recip(C.byref(self.c), _p(pos), tol, _p(f))
        return e, f

    def pme_recip(self, pos, grid, order=5, want_force=False):
        pos = np.ascontiguousarray(pos, np.float64)
        g = np.ascontiguousarray(grid, np.int32)
        f = np.zeros_like(pos) if want_force else None
        e = lib().atm_oracle_pme_recip(C.byref(self.c), _p(pos), _p(g), int(order), _p(f))
        return e, f

    def step(self, params, pos, displ, du_ext=0.0, want_force=True):
        pos = np.ascontiguousarray(pos, np.float64)
        displ = np.ascontiguousarray(displ, np.float64)
        params = np.ascontiguousarray(params, np.float64)
        f = np.zeros_like(pos) if want_force else None
        en = np.zeros(5)
        lib().atm_oracle_step(C.byref(self.c), _p(params), _p(pos), _p(displ), du_ext, _p(f), _p(en))
        return dict(U1=en[0], U2=en[1], u_sc=en[2], energy=en[3], sp=en[4]), f


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


def softcore(u, umax, a, ub):
    fp = C.c_double()
    v = lib().atm_oracle_softcore(u, umax, a, ub, C.byref(fp))
    return v, fp.value


def scalars(params, U1, U2):
    params = np.ascontiguousarray(params, np.float64)
    out = np.zeros(7)
    lib().atm_oracle_scalars(_p(params), C.c_double(U1), C.c_double(U2), _p(out))
    return dict(u_sc=out[0], fp=out[1], ebias=out[2], bfp=out[3], energy=out[4], sp=out[5], sp_ref=out[6])


def displ_table(n, padded, atom_index, dxyz):
    dxyz = np.ascontiguousarray(dxyz, np.float64)
    ai = np.ascontiguousarray(atom_index, np.int32) if atom_index is not None else None
    out = np.empty((padded, 4), np.float32)
    lib().atm_oracle_displ_table(n, padded, _p(ai), _p(dxyz), _p(out))
    return out


def copy_state_f32(posq, corr, displ4):
    posq = np.ascontiguousarray(posq, np.float32)
    displ4 = np.ascontiguousarray(displ4, np.float32)
    n = posq.shape[0]
    p1, p2 = np.empty_like(posq), np.empty_like(posq)
    if corr is not None:
        corr = np.ascontiguousarray(corr, np.float32)
        c1, c2 = np.empty_like(corr), np.empty_like(corr)
    else:
        c1 = c2 = None
    lib().atm_oracle_copy_state_f32(n, _p(posq), _p(corr), _p(displ4), _p(p1), _p(c1), _p(p2), _p(c2))
    return p1, c1, p2, c2


def copy_state_f64(posq, displ4):
    posq = np.ascontiguousarray(posq, np.float64)
    displ4 = np.ascontiguousarray(displ4, np.float32)
    p1, p2 = np.empty_like(posq), np.empty_like(posq)
    lib().atm_oracle_copy_state_f64(posq.shape[0], _p(posq), _p(displ4), _p(p1), _p(p2))
    return p1, p2


def copy_state_ref(pos, displ):
    pos = np.ascontiguousarray(pos, np.float64)
    displ = np.ascontiguousarray(displ, np.float64)
    p1, p2 = np.empty_like(pos), np.empty_like(pos)
    lib().atm_oracle_copy_state_ref(pos.shape[0], _p(pos), _p(displ), _p(p1), _p(p2))
    return p1, p2


def merge_ref(force, f1, f2, sp_ref, direction):
    force = np.array(force, np.float64, copy=True)
    f1 = np.ascontiguousarray(f1, np.float64)
    f2 = np.ascontiguousarray(f2, np.float64)
    lib().atm_oracle_merge_ref(force.shape[0], _p(force), _p(f1), _p(f2), sp_ref, direction)
    return force


def hybrid_force_i64(n, padded, force, f1, f2, sp):
    force = np.array(force, np.int64, copy=True)
    f1 = np.ascontiguousarray(f1, np.int64)
    f2 = np.ascontiguousarray(f2, np.int64)
    lib().atm_oracle_hybrid_force_i64(n, padded, _p(force), _p(f1), _p(f2), sp)
    return force


def num_threads():
    return lib().atm_oracle_num_threads()


def set_num_threads(n):
    lib().atm_oracle_set_num_threads(int(n))
