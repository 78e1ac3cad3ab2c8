"""On-device Hamiltonian replica exchange (atm_hrex_device_*) against the host sweep (atm_hrex_sweep): same decisions,
same parameter rows, no host round trip.  (The reference has no replica exchange)"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _setup(R, seed=7):
    import torch
    import atmmetaforce as atm
    from atmmetaforce import synthetic
    s = synthetic.water_box(6000, n_lig=15, seed=6)
    n = s["pos"].shape[0]
    sched = synthetic.atm_schedule_22()
    be = atm.ATMBackend(n, precision="mixed", num_replicas=R)
    be.set_displacements(s["displ"])
    be.set_box(s["box"])
    be.nb_setup(s["charge"], s["sigma"], s["epsilon"], s["cutoff"], s["ewald_alpha"], skin=0.1, exclusions=s["excl"])
    rng = np.random.default_rng(seed)
    posq = np.zeros((R, be.P, 4), np.float32)
    for r in range(R):
        posq[r, :n, :3] = s["pos"] + rng.normal(0, 0.004, (n, 3)) * (r > 0)
        posq[r, :n, 3] = s["charge"]
    return atm, be, torch.from_numpy(posq).cuda(), sched, s


def test_device_exchange_matches_host_sweep():
    import torch
    atm, be, posq, sched, s = _setup(22)
    R = 22
    rex_d = atm.ReplicaExchange(sched, R, temperature=300.0, seed=11)
    rex_h = atm.ReplicaExchange(sched, R, temperature=300.0, seed=11)
    for k in range(R):
        be.set_parameters(sched[rex_d.replica_state[k]], replica=k)
    stream = torch.cuda.Stream()
    force = torch.zeros((R, 3 * be.P), dtype=torch.int64, device="cuda")
    with torch.cuda.stream(stream):
        be.rebuild(posq, stream=stream)
        rex_d.attach_device(be, stream=stream)
    swaps = 0
    for cycle in range(6):
        with torch.cuda.stream(stream):
            be.step(posq, force, graph=True, stream=stream)
            rex_d.exchange_device(stream=stream)          # asynchronous
        en = be.get_energies(stream=stream)                # (this synchronises: test only)
        changed = rex_h.exchange(en[:, 0:2].copy())
        swaps += len(changed)
        state_d = rex_d.sync_from_device(stream=stream)
        assert np.array_equal(state_d, rex_h.replica_state), cycle
        assert rex_d.accepted == rex_h.accepted
        for k in range(R):   # the device rewrote the parameter rows; the host mirror is refreshed on read
            assert np.array_equal(be.get_parameters(k), sched[rex_h.replica_state[k]])
    assert swaps > 0                                       # the schedule's neighbouring states do exchange here
    assert sorted(rex_d.replica_state.tolist()) == list(range(22))
    # host edits after a device exchange start from the refreshed mirror: only the edited row changes
    row = sched[3].copy(); row[0] += 0.125
    be.set_parameters(row, replica=5)
    with torch.cuda.stream(stream):
        be.step(posq, force, graph=True, stream=stream)
    torch.cuda.synchronize()
    assert np.array_equal(be.get_parameters(5), row)
    assert np.array_equal(be.get_parameters(6), sched[rex_h.replica_state[6]])
    be.close()


def test_device_exchange_flags_non_finite_energy():
    import torch
    atm, be, posq, sched, s = _setup(4)
    rex = atm.ReplicaExchange(sched[:4], 4, temperature=300.0, seed=3)
    for k in range(4):
        be.set_parameters(sched[k], replica=k)
    be.rebuild(posq)
    force = torch.zeros((4, 3 * be.P), dtype=torch.int64, device="cuda")
    be.step(posq, force)
    rex.attach_device(be)
    bad = torch.zeros((4, 2), dtype=torch.float64, device="cuda")
    bad[2, 1] = float("nan")
    be.hrex_exchange(bad, 1)
    with pytest.raises(FloatingPointError):
        rex.sync_from_device()
    assert np.array_equal(rex.replica_state, np.arange(4))   # the cycle was skipped
    be.close()
