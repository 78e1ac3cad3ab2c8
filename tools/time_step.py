"""Developer timing probe (not the bench contract): times atm_step / rebuild for a synthetic system."""
import argparse, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "openmm-atmmetaforce-plugin_b200", "python"))
import numpy as np
import torch
import atmmetaforce as atm
from atmmetaforce import synthetic

ap = argparse.ArgumentParser()
ap.add_argument("--system", default="config3")
ap.add_argument("--replicas", type=int, default=1)
ap.add_argument("--steps", type=int, default=50)
ap.add_argument("--skin", type=float, default=0.1)
ap.add_argument("--natoms", type=int, default=25000)
ap.add_argument("--skin-outer", type=float, default=0.3)
ap.add_argument("--no-energy", action="store_true")
ap.add_argument("--pme", action="store_true")
ap.add_argument("--flush", action="store_true", help="also time the graph step with a 256 MiB L2 flush before every step")
args = ap.parse_args()

s = synthetic.config3() if args.system == "config3" else (synthetic.config4() if args.system == "config4" else synthetic.water_box(args.natoms))
n = s["pos"].shape[0]
R = args.replicas
be = atm.ATMBackend(n, precision="mixed", num_replicas=R)
P = be.P
be.set_displacements(s["displ"])
be.set_box(s["box"])
sched = synthetic.atm_schedule_22()
for r in range(R):
    be.set_parameters(sched[r % 22], replica=r)
be.nb_setup(s["charge"], s["sigma"], s["epsilon"], s["cutoff"], s["ewald_alpha"], skin=args.skin, skin_outer=args.skin_outer, exclusions=s["excl"])
if args.pme:
    grid = synthetic.pme_grid(s["box"], s["ewald_alpha"])
    be.pme_setup(grid)
    print("pme grid", grid)
posq = np.zeros((R, P, 4), np.float32)
rng = np.random.default_rng(0)
for r in range(R):
    posq[r, :n, :3] = s["pos"] + rng.normal(0, 0.002, (n, 3)) * (r > 0)
    posq[r, :n, 3] = s["charge"]
posq = torch.from_numpy(posq).cuda()
force = torch.zeros((R, 3 * P), dtype=torch.int64, device="cuda")
t0 = time.time(); be.rebuild(posq); torch.cuda.synchronize(); t_first = time.time() - t0
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
ev[0].record()
for _ in range(5):
    be.rebuild(posq)
ev[1].record()
be.step(posq, force, collect_stats=True)
en_stats = be.get_energies()
for _ in range(5):
    be.step(posq, force, include_energy=not args.no_energy)
torch.cuda.synchronize()
ev[2].record()
for _ in range(args.steps):
    be.step(posq, force, include_energy=not args.no_energy)
ev[3].record()
evp = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
evp[0].record()
for _ in range(10):
    be.prune(posq)
evp[1].record()
torch.cuda.synchronize()
ms_prune = evp[0].elapsed_time(evp[1]) / 10
en = be.get_energies()
st = be.nb_stats()
if args.flush:
    fl = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    stream = torch.cuda.Stream()
    evs = []
    with torch.cuda.stream(stream):
        for it in range(args.steps + 5):
            fl.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream); be.step(posq, force, include_energy=not args.no_energy, graph=True, stream=stream); b.record(stream)
            if it >= 5:
                evs.append((a, b))
    torch.cuda.synchronize()
    print(f"cold (L2 flushed) graph step: {1e3 * sum(a.elapsed_time(b) for a, b in evs) / len(evs):.1f} us")
ms_step = ev[2].elapsed_time(ev[3]) / args.steps
ms_rebuild = ev[0].elapsed_time(ev[1]) / 5
pairs = en_stats[:, 7].sum()
print(f"system={args.system} N={n} R={R} skin={args.skin} stats={st}")
print(f"first rebuild {t_first*1e3:.1f} ms; rebuild {ms_rebuild:.3f} ms; prune {ms_prune*1e3:.1f} us; step {ms_step*1e3:.1f} us; pairs in cutoff/step {pairs:.3e}; "
      f"listed pair slots {st['list_entries']*8*R:.3e}; fill {pairs/(st['list_entries']*8*R):.3f}; "
      f"Gpairs/s {pairs/ms_step/1e6:.2f}; u[0]={en[0,2]:.3f} U1[0]={en[0,0]:.2f}")
