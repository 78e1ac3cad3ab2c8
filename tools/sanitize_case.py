"""Small two-replica case for compute-sanitizer (memcheck / racecheck / initcheck / synccheck): rebuild, prune, two plain
steps with statistics, one CUDA-graph step (programmatic dependent launches inside), PME on for the last step (tile-owned
shared-memory spread, float transforms, blend, gather), the
on-device replica exchange, energies read back; then the host-buffer pipeline (two handles forked / joined inside one
captured graph) through its rebuild / prune / plain variants.

    compute-sanitizer --tool memcheck  python tools/sanitize_case.py
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "openmm-atmmetaforce-plugin_b200", "python"))
import numpy as np
import torch
import atmmetaforce as atm
from atmmetaforce import synthetic

s = synthetic.water_box(6000, n_lig=15, seed=6)
n = s["pos"].shape[0]
R = 2
sched = synthetic.atm_schedule_22()
be = atm.ATMBackend(n, precision="mixed", num_replicas=R)
be.set_displacements(s["displ"])
be.set_box(s["box"])
for r in range(R):
    be.set_parameters(sched[5 + 11 * r], replica=r)
be.nb_setup(s["charge"], s["sigma"], s["epsilon"], s["cutoff"], s["ewald_alpha"], skin=0.1, skin_outer=0.2, exclusions=s["excl"])
posq = np.zeros((R, be.P, 4), np.float32)
rng = np.random.default_rng(0)
for r in range(R):
    posq[r, :n, :3] = s["pos"] + rng.normal(0, 0.003, (n, 3)) * r
    posq[r, :n, 3] = s["charge"]
posq = torch.from_numpy(posq).cuda()
force = torch.zeros((R, 3 * be.P), dtype=torch.int64, device="cuda")
stream = torch.cuda.Stream()
with torch.cuda.stream(stream):
    be.rebuild(posq, stream=stream)
    be.step(posq, force, collect_stats=True, stream=stream)
    be.prune(posq, stream=stream)
    be.step(posq, force, stream=stream)
    be.step(posq, force, graph=True, stream=stream)
    # round 2: a prune on the side stream concurrent with the step (plain launches, then inside the step's CUDA graph), the
    # copy switch, an asynchronous rebuild (cached graph) behind an odd number of switches, atm_nb_check
    be.step(posq, force, stream=stream, concurrent_prune=True)
    be.step(posq, force, graph=True, stream=stream, concurrent_prune=True)
    be.step(posq, force, graph=True, stream=stream, concurrent_prune=True)
    be.rebuild(posq, stream=stream)
    be.step(posq, force, graph=True, stream=stream)
    rex = atm.ReplicaExchange(sched, R, temperature=300.0, seed=1)
    rex.attach_device(be, stream=stream)   # one rank: atm_hrex_device_cycle without NCCL
    rex.exchange_device(stream=stream)
en = be.get_energies(stream=stream)
be.nb_check(wait=True)
rex.sync_from_device(stream=stream)
be.pme_setup(synthetic.pme_grid(s["box"], s["ewald_alpha"]))
with torch.cuda.stream(stream):
    be.step(posq, force, stream=stream)
en2 = be.get_energies(stream=stream)
print("u:", en[:, 3], "u with PME:", en2[:, 3], "states:", rex.replica_state)
be.close()

# host-buffer pipeline: two one-replica handles, pinned buffers on both sides
bes = []
for r in range(R):
    b = atm.ATMBackend(n, precision="mixed", num_replicas=1)
    b.set_displacements(s["displ"])
    b.set_box(s["box"])
    b.set_parameters(sched[5 + 11 * r])
    b.nb_setup(s["charge"], s["sigma"], s["epsilon"], s["cutoff"], s["ewald_alpha"], skin=0.1, skin_outer=0.2, exclusions=s["excl"])
    bes.append(b)
pipe = atm.HostPipeline(bes)
P = bes[0].P
posq_h = [posq[r:r + 1].cpu().pin_memory() for r in range(R)]
force_h = [torch.zeros((1, 3 * P), dtype=torch.int64).pin_memory() for _ in range(R)]
en_h = [torch.zeros((1, 16), dtype=torch.float64).pin_memory() for _ in range(R)]
for maint in (pipe.REBUILD, pipe.NONE, pipe.PRUNE, pipe.PRUNE_CONCURRENT, pipe.NONE, pipe.REBUILD, pipe.NONE):
    pipe.step(posq_h, force_h, en_h, maintenance=maint, stream=stream)
    stream.synchronize()
pipe.check()
force_f32 = [torch.zeros((1, 3 * P), dtype=torch.float32).pin_memory() for _ in range(R)]
pipe.step(posq_h, force_f32, en_h, maintenance=pipe.NONE, stream=stream)     # float32 force read-back
pipe.step(posq_h, None, en_h, maintenance=pipe.NONE, stream=stream)          # energies only
# external per-state forces / energies of other variable-group forces (atm_host_io.force_state{1,2}_ext_host, energy_ext_host)
f1_ext = [torch.randint(-2 ** 36, 2 ** 36, (1, 3 * P), dtype=torch.int64).pin_memory() for _ in range(R)]
f2_ext = [torch.randint(-2 ** 36, 2 ** 36, (1, 3 * P), dtype=torch.int64).pin_memory() for _ in range(R)]
e_ext = [torch.tensor([[1.5, -2.5]], dtype=torch.float64).pin_memory() for _ in range(R)]
pipe.step(posq_h, force_h, en_h, maintenance=pipe.NONE, stream=stream, f1_ext_host=f1_ext, f2_ext_host=f2_ext, energy_ext_host=e_ext)
stream.synchronize()
print("pipeline u:", [float(e[0, 3]) for e in en_h], "max |F| (fixed point):", [int(f.abs().max()) for f in force_h])
pipe.close()
for b in bes:
    b.close()
