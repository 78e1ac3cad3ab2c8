"""Worker of tests/test_gpu_nccl.py: one rank per GPU (torchrun), the replica layer over REAL NCCL.

Every rank holds its block-cyclic share of 8 replicas of a small water box, steps them, and runs exchange cycles three
ways that must agree bit for bit: (a) atm_hrex_device_cycle on the library's own NCCL communicator (pack -> ncclAllGather
-> sweep, no host round trip), (b) the library's pack / sweep kernels around torch.distributed.all_gather_into_tensor,
(c) the host sweep atm_hrex_sweep on energies gathered through torch.  Prints one line "OK rank=..." per rank."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "openmm-atmmetaforce-plugin_b200", "python"))


def main():
    import torch
    import torch.distributed as dist
    import atmmetaforce as atm
    from atmmetaforce import synthetic
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    NREP = 8
    s = synthetic.water_box(6000, n_lig=15, seed=6)
    n = s["pos"].shape[0]
    sched = synthetic.atm_schedule_22()[:NREP]
    rexs = {k: atm.ReplicaExchange(sched, NREP, rank=rank, world_size=world, temperature=300.0, seed=11) for k in ("library", "torch", "host")}
    mine = rexs["library"].mine
    R = len(mine)
    bes = {}
    rng_pos = {g: s["pos"] + np.random.default_rng(100 + g).normal(0, 0.004, (n, 3)) * (g > 0) for g in mine}
    stream = torch.cuda.Stream(device=dev)
    posq = None
    for k in rexs:
        be = atm.ATMBackend(n, precision="mixed", num_replicas=R, device=local)
        be.set_displacements(s["displ"])
        be.set_box(s["box"])
        be.nb_setup(s["charge"], s["sigma"], s["epsilon"], s["cutoff"], s["ewald_alpha"], skin=0.1, exclusions=s["excl"])
        for j, g in enumerate(mine):
            be.set_parameters(sched[rexs[k].replica_state[g]], replica=j)
        bes[k] = be
        if posq is None:
            p = np.zeros((R, be.P, 4), np.float32)
            for j, g in enumerate(mine):
                p[j, :n, :3] = rng_pos[g]
                p[j, :n, 3] = s["charge"]
            posq = torch.from_numpy(p).to(dev)
        with torch.cuda.stream(stream):
            be.rebuild(posq, stream=stream)
    with torch.cuda.stream(stream):
        rexs["library"].attach_device(bes["library"], stream=stream, collective="library")
        rexs["torch"].attach_device(bes["torch"], stream=stream, collective="torch")
    force = torch.zeros((R, 3 * bes["library"].P), dtype=torch.int64, device=dev)
    swaps = 0
    for cycle in range(6):
        with torch.cuda.stream(stream):
            for k in ("library", "torch", "host"):
                bes[k].step(posq, force, graph=True, stream=stream)
            rexs["library"].exchange_device(stream=stream)
            rexs["torch"].exchange_device(stream=stream)
        en = bes["host"].get_energies(stream=stream)
        u12 = torch.from_numpy(en[:, 0:2].copy()).to(dev)
        for j, row in rexs["host"].exchange(u12):
            bes["host"].set_parameters(row, replica=j)
            swaps += 1
        a = rexs["library"].sync_from_device(stream=stream)
        b = rexs["torch"].sync_from_device(stream=stream)
        h = rexs["host"].replica_state
        assert np.array_equal(a, h) and np.array_equal(b, h), (cycle, a, b, h)
        # every rank holds the same permutation
        got = [None] * world
        dist.all_gather_object(got, a.tolist())
        assert all(x == got[0] for x in got), got
        for j, g in enumerate(mine):
            assert np.array_equal(bes["library"].get_parameters(j), sched[h[g]])
    tot = torch.tensor([swaps], device=dev)
    dist.all_reduce(tot)
    assert int(tot.item()) > 0, "no exchange was accepted: the test does not exercise a swap"
    assert sorted(rexs["library"].replica_state.tolist()) == list(range(NREP))
    print(f"OK rank={rank} world={world} replicas={mine} accepted={rexs['library'].accepted} state={rexs['library'].replica_state.tolist()}", flush=True)
    for be in bes.values():
        be.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
