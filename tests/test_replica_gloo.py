"""world_size-2 (and 3) CPU test of the replica layer over gloo: payload integrity of the all-gather, identical swap
decisions on every rank, the state map stays a permutation, checkpoint round trip."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, os.path.join(ROOT, "openmm-atmmetaforce-plugin_b200", "python"))
    import atmmetaforce as atm
    from atmmetaforce import synthetic
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sched = synthetic.atm_schedule_22()
    rex = atm.ReplicaExchange(sched, 22, rank=rank, world_size=world, seed=5)
    history = []
    for cycle in range(12):
        # every rank fabricates the energies of ITS replicas from the global replica id (so the truth is known)
        local = np.array([[-1e5 - 10.0 * g - cycle, -1e5 - 10.0 * g - cycle + 80.0 + 15.0 * np.sin(g + cycle)] for g in rex.mine])
        gathered = rex.gather(torch.from_numpy(local))
        truth = np.array([[-1e5 - 10.0 * g - cycle, -1e5 - 10.0 * g - cycle + 80.0 + 15.0 * np.sin(g + cycle)] for g in range(22)])
        assert np.array_equal(gathered, truth)
        changed = rex.exchange(torch.from_numpy(local))
        for k, row in changed:
            assert np.array_equal(row, sched[rex.replica_state[rex.mine[k]]])
        history.append(rex.replica_state.copy())
    out[rank] = (np.stack(history), rex.state_dict())
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_replica_exchange_gloo(world):
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29500 + world + (os.getpid() % 200)
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    hist0, sd0 = out[0]
    for r in range(1, world):
        hist, sd = out[r]
        assert np.array_equal(hist, hist0)          # identical decisions on every rank
        assert sd == sd0
    for row in hist0:
        assert sorted(row.tolist()) == list(range(22))
    assert sd0["accepted"] > 0


def test_single_rank_matches_multi_rank_decisions():
    sys.path.insert(0, os.path.join(ROOT, "openmm-atmmetaforce-plugin_b200", "python"))
    import atmmetaforce as atm
    from atmmetaforce import synthetic
    sched = synthetic.atm_schedule_22()
    rex = atm.ReplicaExchange(sched, 22, seed=5)
    for cycle in range(12):
        local = np.array([[-1e5 - 10.0 * g - cycle, -1e5 - 10.0 * g - cycle + 80.0 + 15.0 * np.sin(g + cycle)] for g in range(22)])
        rex.exchange(local)
    sd = rex.state_dict()
    rex2 = atm.ReplicaExchange(sched, 22, seed=1)
    rex2.load_state_dict(sd)
    assert rex2.state_dict() == sd
