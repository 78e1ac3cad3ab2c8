"""ctypes binding of include/atm_b200.h (libatm_b200.so).  No compute happens in Python.

The library is the product path: if it cannot be loaded, or no CUDA device is present, calls fail
loudly (ATMError) -- there is no CPU fallback anywhere in this package.
"""
import ctypes as C
import os

_PKG = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
# ATM_B200_LIB lets a developer A/B an experimental build of the SAME CUDA library; it is not a fallback path
LIB_PATH = os.environ.get("ATM_B200_LIB") or os.path.join(_PKG, "libatm_b200.so")

ATM_OK = 0
PREC_SINGLE, PREC_MIXED, PREC_DOUBLE = 0, 1, 2
NUM_PARAMS = 9
NUM_ENERGY_SLOTS = 16
(E_U1, E_U2, E_U, E_USC, E_EBIAS, E_ENERGY, E_SP, E_NPAIRS, E_NPAIRS_C, E_NPAIRS_S1, E_NPAIRS_S2, E_UREC1, E_UREC2,
 E_USELF) = range(14)
PARAM_NAMES = ("lambda1", "lambda2", "alpha", "u0", "w0", "umax", "ubcore", "acore", "direction")

# every symbol include/atm_b200.h declares (tests check that the library exports all of them)
SYMBOLS = (
    "atm_last_error", "atm_version", "atm_create", "atm_destroy", "atm_set_displacements", "atm_set_parameters",
    "atm_get_parameters", "atm_copy_state", "atm_wrap_positions", "atm_hybrid_force", "atm_softcore_softplus",
    "atm_execute", "atm_get_perturbation_energy", "atm_nb_setup", "atm_pme_setup", "atm_nb_set_dispersion_correction", "atm_set_box", "atm_nb_rebuild", "atm_nb_check", "atm_nb_prune", "atm_step", "atm_step_graph", "atm_profile_enable", "atm_profile_read", "atm_launch_count",
    "atm_energies_device", "atm_get_energies", "atm_nb_stats", "atm_hrex_sweep", "atm_hrex_reduced_energy",
    "atm_hrex_device_setup", "atm_hrex_device_pack", "atm_hrex_device_exchange", "atm_hrex_device_state",
    "atm_re_unique_id", "atm_re_comm_create", "atm_re_comm_from_nccl", "atm_re_comm_destroy", "atm_hrex_device_cycle",
    "atm_host_pipeline_create", "atm_host_pipeline_destroy", "atm_host_pipeline_step", "atm_host_pipeline_check",
    "atm_stream_create", "atm_stream_destroy", "atm_stream_synchronize", "atm_host_alloc", "atm_host_free",
)


class ATMError(Exception):
    """Raised for every non-zero status of the C ABI (the facade maps it the way SWIG maps std::exception,
    reference python/atmmetaforceplugin.i:54-61)."""


class Config(C.Structure):
    _fields_ = [("num_particles", C.c_int32), ("padded_num_particles", C.c_int32), ("precision", C.c_int32),
                ("num_replicas", C.c_int32), ("device", C.c_int32)]


class NonbondedDesc(C.Structure):
    _fields_ = [("charge", C.c_void_p), ("sigma", C.c_void_p), ("epsilon", C.c_void_p),
                ("num_exclusions", C.c_int32), ("exclusions", C.c_void_p),
                ("num_exceptions", C.c_int32), ("exception_pairs", C.c_void_p), ("exception_params", C.c_void_p),
                ("cutoff", C.c_double), ("ewald_alpha", C.c_double), ("skin", C.c_double), ("skin_outer", C.c_double)]


class StepIO(C.Structure):
    _fields_ = [("posq", C.c_void_p), ("posq_corr", C.c_void_p), ("force", C.c_void_p),
                ("force_state1_ext", C.c_void_p), ("force_state2_ext", C.c_void_p), ("energy_ext", C.c_void_p),
                ("posq1", C.c_void_p), ("posq1_corr", C.c_void_p), ("posq2", C.c_void_p), ("posq2_corr", C.c_void_p),
                ("include_energy", C.c_int32), ("collect_stats", C.c_int32), ("concurrent_prune", C.c_int32), ("reserved", C.c_int32)]


class HostIO(C.Structure):
    _fields_ = [("posq_host", C.c_void_p), ("force_host", C.c_void_p), ("energies_host", C.c_void_p),
                ("include_energy", C.c_int32), ("force_format", C.c_int32), ("posq_format", C.c_int32), ("reserved", C.c_int32),
                ("force_state1_ext_host", C.c_void_p), ("force_state2_ext_host", C.c_void_p), ("energy_ext_host", C.c_void_p)]


FORCE_I64, FORCE_F32, FORCE_NONE = 0, 1, 2
POSQ_F4, POSQ_F3 = 0, 1


_lib = None


def lib():
    """Loads the shared library (never builds it implicitly on a GPU box: ship the prebuilt .so)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ATMError(f"{LIB_PATH} is missing: run `python openmm-atmmetaforce-plugin_b200/build.py` "
                       "(the CUDA extension is the product path; there is no fallback)")
    L = C.CDLL(LIB_PATH)
    vp, i32, dbl = C.c_void_p, C.c_int32, C.c_double
    L.atm_last_error.restype = C.c_char_p
    L.atm_version.restype = C.c_char_p
    L.atm_create.argtypes = [C.POINTER(Config), C.POINTER(vp)]
    L.atm_destroy.argtypes = [vp]
    L.atm_set_displacements.argtypes = [vp, vp, vp, vp]
    L.atm_set_parameters.argtypes = [vp, i32, vp]
    L.atm_get_parameters.argtypes = [vp, i32, vp]
    L.atm_copy_state.argtypes = [vp] * 8
    L.atm_wrap_positions.argtypes = [vp] * 5
    L.atm_hybrid_force.argtypes = [vp, vp, vp, vp, dbl, vp]
    L.atm_softcore_softplus.argtypes = [vp, dbl, dbl, vp]
    L.atm_execute.argtypes = [vp, i32, dbl, dbl, vp, vp, vp, i32, C.POINTER(dbl), vp]
    L.atm_get_perturbation_energy.argtypes = [vp, i32, C.POINTER(dbl)]
    L.atm_nb_setup.argtypes = [vp, C.POINTER(NonbondedDesc), vp]
    L.atm_set_box.argtypes = [vp, i32, vp]
    L.atm_pme_setup.argtypes = [vp, i32, i32, i32, i32]
    L.atm_nb_set_dispersion_correction.argtypes = [vp, i32]
    L.atm_nb_rebuild.argtypes = [vp, vp, vp]
    L.atm_nb_prune.argtypes = [vp, vp, vp]
    L.atm_nb_check.argtypes = [vp, i32]
    L.atm_step.argtypes = [vp, C.POINTER(StepIO), vp]
    L.atm_step_graph.argtypes = [vp, C.POINTER(StepIO), vp]
    L.atm_profile_enable.argtypes = [vp, i32]
    L.atm_profile_read.argtypes = [vp, C.POINTER(dbl), C.POINTER(i32)]
    L.atm_launch_count.argtypes = [vp, C.POINTER(C.c_uint64)]
    L.atm_energies_device.argtypes = [vp, C.POINTER(vp)]
    L.atm_get_energies.argtypes = [vp, vp, vp]
    L.atm_nb_stats.argtypes = [vp, vp]
    L.atm_hrex_sweep.argtypes = [i32, vp, i32, vp, vp, dbl, C.c_uint64, C.c_uint64, C.POINTER(i32)]
    L.atm_hrex_reduced_energy.argtypes = [vp, dbl, dbl, dbl]
    L.atm_hrex_device_setup.argtypes = [vp, i32, vp, i32, vp, vp, vp, i32, dbl, C.c_uint64, vp]
    L.atm_hrex_device_pack.argtypes = [vp, vp, i32, vp]
    L.atm_hrex_device_exchange.argtypes = [vp, vp, C.c_uint64, vp]
    L.atm_hrex_device_state.argtypes = [vp, vp, vp, vp]
    L.atm_hrex_reduced_energy.restype = dbl
    L.atm_host_pipeline_create.argtypes = [i32, C.POINTER(vp), C.POINTER(vp)]
    L.atm_host_pipeline_destroy.argtypes = [vp]
    L.atm_re_unique_id.argtypes = [vp]
    L.atm_re_comm_create.argtypes = [vp, i32, i32, i32, C.POINTER(vp)]
    L.atm_re_comm_from_nccl.argtypes = [vp, i32, C.POINTER(vp)]
    L.atm_re_comm_destroy.argtypes = [vp]
    L.atm_hrex_device_cycle.argtypes = [vp, vp, C.c_uint64, vp]
    L.atm_host_pipeline_step.argtypes = [vp, C.POINTER(HostIO), i32, vp]
    L.atm_host_pipeline_check.argtypes = [vp]
    L.atm_stream_create.argtypes = [i32, C.POINTER(vp)]
    L.atm_stream_destroy.argtypes = [vp]
    L.atm_stream_synchronize.argtypes = [vp]
    L.atm_host_alloc.argtypes = [C.c_size_t, C.POINTER(vp)]
    L.atm_host_free.argtypes = [vp]
    for name in SYMBOLS:
        fn = getattr(L, name)
        if fn.restype is C.c_int and name not in ("atm_last_error", "atm_version", "atm_hrex_reduced_energy"):
            fn.restype = C.c_int
    _lib = L
    return L


def check(rc):
    if rc != ATM_OK:
        raise ATMError(lib().atm_last_error().decode() or f"atm_b200 status {rc}")
