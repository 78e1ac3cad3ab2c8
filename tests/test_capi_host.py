"""CPU tests of the product's host side: the shared library loads, exports every symbol include/atm_b200.h
declares, its host-only entry points agree with the oracle, and device entry points fail loudly without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import atmmetaforce as atm
from atmmetaforce import _capi
import oracle_py as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "atm_b200.h")).read()
    declared = set(re.findall(r"\b(atm_[a-z0-9_]+)\s*\(", header))
    declared -= {"atm_handle"}
    lib = C.CDLL(_capi.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/atm_b200.h but not exported"
    assert declared == set(_capi.SYMBOLS), declared ^ set(_capi.SYMBOLS)
    assert _capi.lib().atm_version().decode().startswith("0.3.1")


def test_create_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(atm.ATMError) as e:
        atm.ATMBackend(100)
    assert "no CUDA device" in str(e.value) or "CUDA" in str(e.value)


def test_host_scalar_stage_matches_oracle():
    rng = np.random.default_rng(0)
    for _ in range(300):
        p = [rng.uniform(0, 1), rng.uniform(0, 1), rng.choice([0.0, rng.uniform(0.001, 0.3)]), rng.uniform(-50, 500),
             rng.uniform(-5, 5), 836.8, 418.4, 0.0625, rng.choice([1.0, -1.0])]
        U1 = rng.uniform(-2e5, 0)
        U2 = U1 + rng.uniform(-300, 3000)
        a = atm.softcore_softplus(p, U1, U2)
        b = O.scalars(p, U1, U2)
        for k in ("u_sc", "fp", "ebias", "bfp", "energy", "sp", "sp_ref"):
            if np.isinf(b[k]) or np.isnan(b[k]):     # exp overflow quirk: both sides must agree
                assert (np.isinf(a[k]) and a[k] == b[k]) or (np.isnan(a[k]) and np.isnan(b[k]))
            else:
                assert abs(a[k] - b[k]) <= 1e-12 * max(1.0, abs(b[k])), (k, a[k], b[k])


def test_softplus_overflow_quirk_is_preserved():
    """Very negative alpha (u_sc - u0): exp overflows to inf exactly as in the reference."""
    p = [0.0, 0.5, 1.0, 1e6, 0.0, 1e9, 5e8, 0.0625, 1.0]
    a = atm.softcore_softplus(p, 0.0, 10.0)
    b = O.scalars(p, 0.0, 10.0)
    assert np.isinf(a["ebias"]) and np.isinf(b["ebias"]) and a["bfp"] == b["bfp"] == 0.0


def test_bad_arguments_raise():
    with pytest.raises(atm.ATMError):
        atm.hrex_sweep(np.zeros((2, 9)), np.zeros((2, 2)), [0, 5], 0.4, 1, 0)   # invalid state index
    with pytest.raises(atm.ATMError):
        atm.hrex_sweep(np.zeros((2, 9)), np.array([[0.0, np.nan], [0, 0]]), [0, 1], 0.4, 1, 0)  # NaN guard


def test_host_pipeline_rejects_bad_arguments_before_touching_the_device():
    """atm_host_pipeline_*: argument checks come first, so they are testable without a GPU (no CPU fallback exists:
    a valid call needs handles, and atm_create needs a CUDA device)."""
    L = _capi.lib()
    p = C.c_void_p()
    assert L.atm_host_pipeline_create(0, None, C.byref(p)) != _capi.ATM_OK and not p.value
    assert b"no handles" in L.atm_last_error()
    handles = (C.c_void_p * 1)(None)
    assert L.atm_host_pipeline_create(1, handles, C.byref(p)) != _capi.ATM_OK and not p.value
    assert L.atm_host_pipeline_step(None, None, 0, None) != _capi.ATM_OK
    assert L.atm_host_pipeline_destroy(None) == _capi.ATM_OK      # destroying nothing is not an error
    assert C.sizeof(_capi.HostIO) == 64                           # 3 pointers + 4 int32 + 3 pointers, as declared in the header


# ------------------------------------------------------------------ replica exchange decisions (host, deterministic)

def _schedule():
    from atmmetaforce import synthetic
    return synthetic.atm_schedule_22()


def test_hrex_is_deterministic_and_a_permutation():
    sched = _schedule()
    rng = np.random.default_rng(3)
    state = np.arange(22, dtype=np.int32)
    for cycle in range(50):
        u12 = np.stack([rng.uniform(-1e5, -9e4, 22), np.zeros(22)], 1)
        u12[:, 1] = u12[:, 0] + rng.normal(100, 80, 22)
        a, na = atm.hrex_sweep(sched, u12, state, 0.4, 7, cycle)
        b, nb = atm.hrex_sweep(sched, u12, state, 0.4, 7, cycle)
        assert np.array_equal(a, b) and na == nb                      # every rank takes the same decision
        assert sorted(a.tolist()) == list(range(22))                 # still a permutation
        state = a
    assert not np.array_equal(state, np.arange(22))                   # something was exchanged


def test_hrex_reduced_energy_matches_oracle():
    sched = _schedule()
    for p in sched:
        e = atm.hrex_reduced_energy(p, -1000.0, -950.0, 0.4)
        assert abs(e - 0.4 * O.scalars(p, -1000.0, -950.0)["energy"]) < 1e-9


def test_hrex_detailed_balance_two_states():
    """Two states, two replicas with fixed energies: acceptance frequency -> min(1, exp(-delta))."""
    sched = _schedule()[[2, 3]]
    u12 = np.array([[-100.0, -60.0], [-100.0, -20.0]])
    def red(p, u):
        return atm.hrex_reduced_energy(p, u[0], u[1], 0.4)
    delta = (red(sched[0], u12[1]) + red(sched[1], u12[0])) - (red(sched[0], u12[0]) + red(sched[1], u12[1]))
    expect = min(1.0, float(np.exp(-delta)))
    acc = 0
    trials = 4000
    for c in range(trials):
        _, n = atm.hrex_sweep(sched, u12, [0, 1], 0.4, 99, 2 * c)   # even cycles pair states (0,1)
        acc += n
    assert abs(acc / trials - expect) < 0.03, (acc / trials, expect)


def test_partition_replicas_block_cyclic():
    from atmmetaforce import synthetic
    for w in (1, 2, 4, 8):
        parts = synthetic.partition_replicas(22, w)
        assert sorted(sum(parts, [])) == list(range(22))
        assert max(map(len, parts)) - min(map(len, parts)) <= 1
    assert [len(p) for p in synthetic.partition_replicas(22, 8)] == [3, 3, 3, 3, 3, 3, 2, 2]


def test_synthetic_systems_are_well_formed():
    from atmmetaforce import synthetic
    s = synthetic.config3()
    n = s["pos"].shape[0]
    assert 22000 < n < 26000 and abs(s["charge"].sum()) < 1e-9
    assert s["lig1"].size == 40 and np.count_nonzero(np.abs(s["displ"]).sum(1)) == 40
    assert s["excl"].min() >= 0 and s["excl"].max() < n and (s["excl"][:, 0] != s["excl"][:, 1]).all()
    assert abs(s["ewald_alpha"] - 2.9203) < 1e-3
    sched = synthetic.atm_schedule_22()
    assert sched.shape == (22, 9) and (sched[:11, 8] == 1).all() and (sched[11:, 8] == -1).all()
