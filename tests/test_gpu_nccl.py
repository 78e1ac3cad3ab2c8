"""The exchange cycle of the replica layer over REAL NCCL with more than one rank: spawns one process per
GPU with torchrun; skipped on a one-GPU box.  tests/test_replica_gloo.py covers the same host logic on CPU with gloo."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_nccl_cycle_matches_torch_collective_and_host_sweep():
    import torch
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs at least two GPUs on the box (run: gpurun --gpus 2 -- python -m pytest tests/test_gpu_nccl.py -m gpu)")
    world = 2 if ngpu < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29517", os.path.join(ROOT, "tests", "nccl_hrex_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    import re
    ok = re.findall(r"OK rank=(\d+) world=\d+ .*?state=(\[[^\]]*\])", out.stdout)   # the ranks' lines may interleave
    assert sorted(int(r) for r, _ in ok) == list(range(world)), out.stdout[-3000:]
    assert len({st for _, st in ok}) == 1          # identical replica_state on every rank


def test_one_rank_cycle_needs_no_nccl():
    """atm_hrex_device_cycle with comm = NULL equals pack + exchange."""
    import numpy as np
    import torch
    import atmmetaforce as atm
    from atmmetaforce import synthetic
    s = synthetic.water_box(6000, n_lig=15, seed=6)
    n = s["pos"].shape[0]
    sched = synthetic.atm_schedule_22()[:4]
    states = []
    for mode in ("library", "torch"):
        be = atm.ATMBackend(n, precision="mixed", num_replicas=4)
        be.set_displacements(s["displ"])
        be.set_box(s["box"])
        be.nb_setup(s["charge"], s["sigma"], s["epsilon"], s["cutoff"], s["ewald_alpha"], skin=0.1, exclusions=s["excl"])
        rex = atm.ReplicaExchange(sched, 4, temperature=300.0, seed=5)
        for k in range(4):
            be.set_parameters(sched[k], replica=k)
        posq = np.zeros((4, be.P, 4), np.float32)
        rng = np.random.default_rng(1)
        for r in range(4):
            posq[r, :n, :3] = s["pos"] + rng.normal(0, 0.004, (n, 3)) * (r > 0)
            posq[r, :n, 3] = s["charge"]
        posq = torch.from_numpy(posq).cuda()
        force = torch.zeros((4, 3 * be.P), dtype=torch.int64, device="cuda")
        be.rebuild(posq)
        rex.attach_device(be, collective=mode)
        for _ in range(5):
            be.step(posq, force)
            rex.exchange_device()
        states.append(rex.sync_from_device().tolist())
        sd = rex.state_dict()
        assert sd["replica_state"] == states[-1]
        be.close()
    assert states[0] == states[1]
