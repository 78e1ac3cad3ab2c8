"""Thin object wrapper over the C ABI that takes torch CUDA tensors as borrowed device buffers.

torch is plumbing only (device memory + the current stream); every kernel is in libatm_b200.so.
Mirrors the reference's kernel seam CalcATMMetaForceKernel (openmmapi/include/ATMMetaForceKernels.h:15-59):
initialize -> __init__/set_displacements, copyState -> copy_state, execute -> execute,
copyParametersToContext -> set_displacements, getPerturbationEnergy -> get_perturbation_energy.
"""
import ctypes as C

import numpy as np

from . import _capi
from ._capi import ATMError, check

_PREC = {"single": _capi.PREC_SINGLE, "mixed": _capi.PREC_MIXED, "double": _capi.PREC_DOUBLE}


def _stream_ptr(stream=None):
    import torch
    s = stream if stream is not None else torch.cuda.current_stream()
    return C.c_void_p(s.cuda_stream)


def _dptr(t):
    if t is None:
        return None
    if not t.is_cuda:
        raise ATMError("expected a CUDA tensor (this back-end has no CPU path)")
    if not t.is_contiguous():
        raise ATMError("expected a contiguous tensor")
    return C.c_void_p(t.data_ptr())


def _np(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


class ATMBackend:
    """One handle == one OpenMM Context (or R batched replicas of one System)."""

    def __init__(self, num_particles, padded_num_particles=0, precision="mixed", num_replicas=1, device=-1):
        self._h = C.c_void_p()
        cfg = _capi.Config(int(num_particles), int(padded_num_particles), _PREC[precision], int(num_replicas), int(device))
        check(_capi.lib().atm_create(C.byref(cfg), C.byref(self._h)))
        self.N = int(num_particles)
        self.P = int(padded_num_particles) if padded_num_particles else 32 * ((self.N + 31) // 32)
        self.R = int(num_replicas)
        self.precision = precision
        self._keep = []

    def close(self):
        if self._h:
            _capi.lib().atm_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- displacement table / parameters ---------------------------------------------------------
    def set_displacements(self, dxyz, atom_index=None, stream=None):
        d = _np(dxyz, np.float64).reshape(self.N, 3)
        ai = _np(atom_index, np.int32) if atom_index is not None else None
        check(_capi.lib().atm_set_displacements(self._h, ai.ctypes.data_as(C.c_void_p) if ai is not None else None,
                                                d.ctypes.data_as(C.c_void_p), _stream_ptr(stream)))

    def set_parameters(self, p, replica=-1):
        p = _np(p, np.float64)
        assert p.size == _capi.NUM_PARAMS
        check(_capi.lib().atm_set_parameters(self._h, int(replica), p.ctypes.data_as(C.c_void_p)))

    def get_parameters(self, replica=0):
        p = np.zeros(_capi.NUM_PARAMS)
        check(_capi.lib().atm_get_parameters(self._h, int(replica), p.ctypes.data_as(C.c_void_p)))
        return p

    # -- Tier 1 -------------------------------------------------------------------------------------
    def copy_state(self, posq, posq1, posq2, posq_corr=None, posq1_corr=None, posq2_corr=None, stream=None):
        check(_capi.lib().atm_copy_state(self._h, _dptr(posq), _dptr(posq_corr), _dptr(posq1), _dptr(posq1_corr),
                                         _dptr(posq2), _dptr(posq2_corr), _stream_ptr(stream)))

    def wrap_positions(self, posq_in, posq_out, box, stream=None):
        b = _np(box, np.float64).reshape(9)
        check(_capi.lib().atm_wrap_positions(self._h, _dptr(posq_in), _dptr(posq_out), b.ctypes.data_as(C.c_void_p),
                                             _stream_ptr(stream)))

    def hybrid_force(self, force, f1, f2, sp, stream=None):
        check(_capi.lib().atm_hybrid_force(self._h, _dptr(force), _dptr(f1), _dptr(f2), float(sp), _stream_ptr(stream)))

    def execute(self, U1, U2, force, f1, f2, include_energy=True, replica=0, stream=None):
        e = C.c_double()
        check(_capi.lib().atm_execute(self._h, int(replica), float(U1), float(U2), _dptr(force), _dptr(f1), _dptr(f2),
                                      1 if include_energy else 0, C.byref(e), _stream_ptr(stream)))
        return e.value

    def get_perturbation_energy(self, replica=0):
        u = C.c_double()
        check(_capi.lib().atm_get_perturbation_energy(self._h, int(replica), C.byref(u)))
        return u.value

    # -- Tier 2 -------------------------------------------------------------------------------------
    def nb_setup(self, charge, sigma, epsilon, cutoff, ewald_alpha, skin=0.1, skin_outer=0.0, exclusions=None, exception_pairs=None,
                 exception_params=None, stream=None):
        q, s, e = _np(charge, np.float64), _np(sigma, np.float64), _np(epsilon, np.float64)
        assert q.size == self.N and s.size == self.N and e.size == self.N
        ex = _np(exclusions if exclusions is not None else np.zeros((0, 2)), np.int32).reshape(-1, 2)
        xp = _np(exception_pairs if exception_pairs is not None else np.zeros((0, 2)), np.int32).reshape(-1, 2)
        xq = _np(exception_params if exception_params is not None else np.zeros((0, 3)), np.float64).reshape(-1, 3)
        assert xp.shape[0] == xq.shape[0]
        d = _capi.NonbondedDesc(q.ctypes.data, s.ctypes.data, e.ctypes.data, ex.shape[0], ex.ctypes.data, xp.shape[0],
                                xp.ctypes.data, xq.ctypes.data, float(cutoff), float(ewald_alpha), float(skin), float(skin_outer))
        check(_capi.lib().atm_nb_setup(self._h, C.byref(d), _stream_ptr(stream)))

    def pme_setup(self, grid, order=5):
        """Switch on the two-state PME reciprocal space (grid = (nx, ny, nz); (0,0,0) switches it off)."""
        check(_capi.lib().atm_pme_setup(self._h, int(grid[0]), int(grid[1]), int(grid[2]), int(order)))

    def set_dispersion_correction(self, on=True):
        """NonbondedForce.setUseDispersionCorrection: add the long-range LJ tail energy (a constant / V) to U1 and U2."""
        check(_capi.lib().atm_nb_set_dispersion_correction(self._h, 1 if on else 0))

    def set_box(self, box, replica=-1):
        b = np.asarray(box, np.float64)
        if b.size == 3:
            b = np.diag(b)
        b = _np(b, np.float64).reshape(9)
        check(_capi.lib().atm_set_box(self._h, int(replica), b.ctypes.data_as(C.c_void_p)))

    def rebuild(self, posq, stream=None):
        check(_capi.lib().atm_nb_rebuild(self._h, _dptr(posq), _stream_ptr(stream)))

    def nb_check(self, wait=True):
        """Verification of the last asynchronous rebuild; raises ATMError when a pair list outgrew its capacity (the
        steps since then returned NaN: rebuild again and repeat them)."""
        check(_capi.lib().atm_nb_check(self._h, 1 if wait else 0))

    def prune(self, posq, stream=None):
        check(_capi.lib().atm_nb_prune(self._h, _dptr(posq), _stream_ptr(stream)))

    def step(self, posq, force, posq_corr=None, f1_ext=None, f2_ext=None, energy_ext=None, posq1=None, posq1_corr=None,
             posq2=None, posq2_corr=None, include_energy=True, collect_stats=False, graph=False, stream=None, concurrent_prune=False):
        """atm_step / atm_step_graph.  concurrent_prune: re-prune the inner pair list from these coordinates on the
        back-end's side stream while the step runs; the next step uses the new list."""
        def v(t):
            p = _dptr(t)
            return p.value if p is not None else None
        io = _capi.StepIO(v(posq), v(posq_corr), v(force), v(f1_ext), v(f2_ext), v(energy_ext), v(posq1), v(posq1_corr),
                          v(posq2), v(posq2_corr), 1 if include_energy else 0, 1 if collect_stats else 0,
                          1 if concurrent_prune else 0, 0)
        fn = _capi.lib().atm_step_graph if graph else _capi.lib().atm_step
        check(fn(self._h, C.byref(io), _stream_ptr(stream)))

    # ---- on-device Hamiltonian replica exchange (no host round trip; include/atm_b200.h atm_hrex_device_*)
    def hrex_setup(self, schedule, replica_state, local_replicas, gather_slot, gathered_rows, beta, seed, stream=None):
        sp = _np(schedule, np.float64).reshape(-1, _capi.NUM_PARAMS)
        rs = _np(replica_state, np.int32)
        loc = np.full(self.R, -1, np.int32)
        loc[:len(local_replicas)] = np.asarray(local_replicas, np.int32)
        gs = _np(gather_slot, np.int32)
        check(_capi.lib().atm_hrex_device_setup(self._h, sp.shape[0], sp.ctypes.data_as(C.c_void_p), rs.size,
                                                rs.ctypes.data_as(C.c_void_p), loc.ctypes.data_as(C.c_void_p),
                                                gs.ctypes.data_as(C.c_void_p), int(gathered_rows), float(beta), int(seed),
                                                _stream_ptr(stream)))

    def hrex_pack(self, send, stream=None):
        """send: [rows][2] float64 CUDA tensor <- (U1, U2) of the local replicas (rows >= R zero-filled)."""
        check(_capi.lib().atm_hrex_device_pack(self._h, _dptr(send), int(send.shape[0]), _stream_ptr(stream)))

    def hrex_exchange(self, gathered, cycle, stream=None):
        check(_capi.lib().atm_hrex_device_exchange(self._h, _dptr(gathered), int(cycle), _stream_ptr(stream)))

    def hrex_cycle(self, comm, cycle, stream=None):
        """atm_hrex_device_cycle: pack -> ncclAllGather (inside the library) -> sweep, on `stream`.  comm: ReplicaComm or None."""
        check(_capi.lib().atm_hrex_device_cycle(self._h, comm._c if comm is not None else None, int(cycle), _stream_ptr(stream)))

    def hrex_state(self, num_replicas, stream=None):
        """Synchronises.  Returns (replica_state[num_replicas], accepted swaps, cycles, error flag)."""
        rs = np.zeros(int(num_replicas), np.int32)
        cnt = np.zeros(3, np.int64)
        check(_capi.lib().atm_hrex_device_state(self._h, rs.ctypes.data_as(C.c_void_p), cnt.ctypes.data_as(C.c_void_p),
                                                _stream_ptr(stream)))
        return rs, int(cnt[0]), int(cnt[1]), int(cnt[2])

    def get_parameters(self, replica=0):
        p = np.zeros(_capi.NUM_PARAMS)
        check(_capi.lib().atm_get_parameters(self._h, int(replica), p.ctypes.data_as(C.c_void_p)))
        return p

    def profile_enable(self, on=True):
        check(_capi.lib().atm_profile_enable(self._h, 1 if on else 0))

    def profile_read(self):
        """(total nb2 milliseconds, number of nb2 launches) since the last read; synchronises the recorded events."""
        ms, n = C.c_double(), C.c_int32()
        check(_capi.lib().atm_profile_read(self._h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def launch_count(self):
        n = C.c_uint64()
        check(_capi.lib().atm_launch_count(self._h, C.byref(n)))
        return n.value

    def energies_device_ptr(self):
        p = C.c_void_p()
        check(_capi.lib().atm_energies_device(self._h, C.byref(p)))
        return p.value

    def get_energies(self, stream=None):
        """Synchronises.  Returns an [R][8] array (see _capi.E_* slots)."""
        out = np.zeros((self.R, _capi.NUM_ENERGY_SLOTS))
        check(_capi.lib().atm_get_energies(self._h, out.ctypes.data_as(C.c_void_p), _stream_ptr(stream)))
        return out

    def nb_stats(self):
        out = np.zeros(8, np.int64)
        check(_capi.lib().atm_nb_stats(self._h, out.ctypes.data_as(C.c_void_p)))
        keys = ("sites", "clusters", "list_entries", "cap_env", "cap_lig", "displaced_atoms", "groups", "columns")
        return dict(zip(keys, (int(x) for x in out)))


class ReplicaComm:
    """atm_re_comm_*: the NCCL communicator of the replica layer, owned by the library.  `share` is any callable that
    takes the 128-byte id from rank 0 and returns it on every rank (e.g. a torch.distributed broadcast)."""

    def __init__(self, rank, world_size, device, share=None):
        self.rank, self.world = int(rank), int(world_size)
        self._c = C.c_void_p()
        idbuf = (C.c_char * 128)()
        if self.world > 1:
            if share is None:
                raise ATMError("ReplicaComm: a `share` callable is needed to distribute the NCCL unique id")
            raw = None
            if self.rank == 0:   # a failure here (no usable libnccl) is shared too, so that no rank waits for an id that never comes
                raw = bytes(idbuf.raw) if _capi.lib().atm_re_unique_id(idbuf) == _capi.ATM_OK else b""
                if raw:
                    raw = bytes(idbuf.raw)
            raw = share(raw)
            if not raw:
                raise ATMError("ReplicaComm: rank 0 could not create an NCCL unique id: " + (_capi.lib().atm_last_error().decode() or "?"))
            idbuf = (C.c_char * 128).from_buffer_copy(raw)
        check(_capi.lib().atm_re_comm_create(idbuf, self.world, self.rank, int(device), C.byref(self._c)))

    def close(self):
        if self._c:
            _capi.lib().atm_re_comm_destroy(self._c)
            self._c = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class HostPipeline:
    """atm_host_pipeline_* (include/atm_b200.h): the step with pinned HOST buffers on both sides, for one or more
    back-ends ("chunks") whose copies and kernels overlap; one cached CUDA graph launch per step.

    posq_host  : list of pinned CPU float32 tensors [R_c][P][4] (coordinates + charge, OpenMM's posq layout) or [R_c][P][3]
                 (packed coordinates, 12 B per slot: the charges come from nb_setup); the last dimension selects the format
    force_host : list of pinned CPU tensors [R_c][3P]: int64 (2^32 fixed point, OpenMM's long force layout) or float32
                 (kJ/mol/nm, half the D2H bytes); the dtype selects the format.  None = energies only
    energies_host : list of pinned CPU float64 tensors [R_c][NUM_ENERGY_SLOTS] or None
    """
    NONE, PRUNE, REBUILD, PRUNE_CONCURRENT = 0, 1, 2, 3

    def __init__(self, backends):
        self.backends = list(backends)
        arr = (C.c_void_p * len(self.backends))(*[b._h for b in self.backends])
        self._p = C.c_void_p()
        check(_capi.lib().atm_host_pipeline_create(len(self.backends), arr, C.byref(self._p)))
        self._ios = None
        self._key = None

    def close(self):
        if self._p:
            _capi.lib().atm_host_pipeline_destroy(self._p)
            self._p = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @staticmethod
    def _hptr(t, what):
        if t is None:
            return None
        if t.is_cuda or not t.is_contiguous() or not t.is_pinned():
            raise ATMError(f"{what}: expected a contiguous pinned host tensor")
        return t.data_ptr()

    def step(self, posq_host, force_host, energies_host=None, maintenance=0, include_energy=True, stream=None,
             f1_ext_host=None, f2_ext_host=None, energy_ext_host=None):
        """f1_ext_host / f2_ext_host ([R][3P] int64 fixed point) and energy_ext_host ([R][2] float64), one pinned tensor
        per back-end: contributions of other variable-group forces evaluated by the caller at the state-1 / state-2
        coordinates (atm_host_io.force_state{1,2}_ext_host / energy_ext_host)."""
        n = len(self.backends)
        import torch
        if force_host is None:
            force_host = [None] * n
        if len(posq_host) != n or len(force_host) != n or (energies_host is not None and len(energies_host) != n):
            raise ATMError("HostPipeline.step: one buffer per back-end is required")
        ext = (f1_ext_host, f2_ext_host, energy_ext_host)
        for e, dt, per in zip(ext, (torch.int64, torch.int64, torch.float64), (None, None, 2)):
            if e is None:
                continue
            if len(e) != n:
                raise ATMError("HostPipeline.step: one external buffer per back-end is required")
            for c, b in enumerate(self.backends):
                if e[c].dtype != dt or e[c].numel() != (b.R * per if per else b.R * b.P * 3):
                    raise ATMError(f"HostPipeline.step: chunk {c}: external buffer has the wrong dtype or size")
        key = (tuple(t.data_ptr() for t in posq_host), tuple(t.data_ptr() if t is not None else 0 for t in force_host),
               tuple(t.data_ptr() for t in energies_host) if energies_host is not None else None, bool(include_energy),
               tuple(tuple(t.data_ptr() for t in e) if e is not None else None for e in ext))
        if key != self._key:
            ios = (_capi.HostIO * n)()
            for c, b in enumerate(self.backends):
                f = force_host[c]
                if posq_host[c].numel() == b.R * b.P * 4:
                    pfmt = _capi.POSQ_F4
                elif posq_host[c].numel() == b.R * b.P * 3:
                    pfmt = _capi.POSQ_F3
                else:
                    raise ATMError(f"HostPipeline.step: chunk {c}: posq_host size does not match [R][P][4] or [R][P][3]")
                if f is not None and f.numel() != b.R * b.P * 3:
                    raise ATMError(f"HostPipeline.step: chunk {c}: buffer size does not match [R][P]")
                if f is None:
                    fmt = _capi.FORCE_NONE
                elif f.dtype == torch.int64:
                    fmt = _capi.FORCE_I64
                elif f.dtype == torch.float32:
                    fmt = _capi.FORCE_F32
                else:
                    raise ATMError("HostPipeline.step: force_host must be int64 (fixed point) or float32")
                ios[c] = _capi.HostIO(self._hptr(posq_host[c], "posq_host"), self._hptr(f, "force_host"),
                                      self._hptr(energies_host[c], "energies_host") if energies_host is not None else None,
                                      1 if include_energy else 0, fmt, pfmt, 0,
                                      *[self._hptr(e[c], "external buffer") if e is not None else None for e in ext])
            self._ios, self._key = ios, key
        check(_capi.lib().atm_host_pipeline_step(self._p, self._ios, int(maintenance), _stream_ptr(stream)))

    def check(self):
        """atm_host_pipeline_check: after synchronising a REBUILD step; raises ATMError if that rebuild overflowed."""
        check(_capi.lib().atm_host_pipeline_check(self._p))


def softcore_softplus(params, U1, U2):
    """Host scalar stage of the library (same math the device runs)."""
    p = _np(params, np.float64)
    out = np.zeros(7)
    check(_capi.lib().atm_softcore_softplus(p.ctypes.data_as(C.c_void_p), float(U1), float(U2), out.ctypes.data_as(C.c_void_p)))
    keys = ("u_sc", "fp", "ebias", "bfp", "energy", "sp", "sp_ref")
    return dict(zip(keys, out))


def hrex_sweep(state_params, u12, replica_state, beta, seed, cycle):
    sp = _np(state_params, np.float64).reshape(-1, _capi.NUM_PARAMS)
    u = _np(u12, np.float64).reshape(-1, 2)
    rs = np.array(replica_state, dtype=np.int32, copy=True)
    acc = C.c_int32()
    check(_capi.lib().atm_hrex_sweep(sp.shape[0], sp.ctypes.data_as(C.c_void_p), u.shape[0], u.ctypes.data_as(C.c_void_p),
                                     rs.ctypes.data_as(C.c_void_p), float(beta), int(seed), int(cycle), C.byref(acc)))
    return rs, acc.value


def hrex_reduced_energy(params, U1, U2, beta):
    p = _np(params, np.float64)
    return _capi.lib().atm_hrex_reduced_energy(p.ctypes.data_as(C.c_void_p), float(U1), float(U2), float(beta))
