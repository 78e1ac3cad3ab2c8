"""Replica-exchange driver: runs a lambda schedule end to end on the GPUs of one node and writes the on-disk formats
the reference's example scripts produce (the reference itself has no driver -- one script
runs one lambda state, ref: example/abfe/abfe.py:94-160 -- this is the caller layer an AToM-style workflow puts on top).

One process per GPU.  Every process owns its block-cyclic share of the replicas, batched in ONE back-end handle; per
cycle it advances them `steps_per_cycle` steps (pair-list prune / rebuild on their cadences), appends one sample line
per replica to `<job>_r<replica>.out` (ref format: README.md:193-201), and runs one Hamiltonian exchange cycle on the
device (pack -> NCCL all-gather -> sweep; coordinates never move between GPUs, lambda states do).  Every
`checkpoint_every` cycles each replica's State XML (`<job>_r<replica>-chk.xml`, the format the examples load and save,
ref: example/abfe/temoa-g1-equil.xml) and the exchange bookkeeping (`<job>-hrex.json`) are rewritten; restart() picks
them up.

The integrator is a seam, not part of the ATM hot path: `propagate(step, posq, force)` receives the device coordinates
[R][P][4] float32 and the merged ATM force [R][3P] int64 (2^32 fixed point) of that step and updates posq in place.  In
production that is OpenMM's integrator (with the bonded / reciprocal-space forces it owns); the default here moves the
atoms by seeded Gaussian noise around their start (what bench.py does), enough to exercise every file and cadence.
"""
import json
import os

import numpy as np

from . import io
from .backend import ATMBackend
from .replica import ReplicaExchange

PARAM_NAMES = ("ATMLambda1", "ATMLambda2", "ATMAlpha", "ATMU0", "ATMW0", "ATMUmax", "ATMUbcore", "ATMAcore", "ATMDirection")


class JitterPropagator:
    """x(step) = x(0) + clipped N(0, sigma) per atom and step, seeded (no dynamics: a stand-in for the integrator)."""

    def __init__(self, sigma_nm=0.005, seed=2022, fields=8):
        self.sigma, self.seed, self.fields = float(sigma_nm), int(seed), int(fields)
        self._base = self._noise = None

    def set_base(self, posq):
        """The coordinates the noise is added to (the driver passes the System's start coordinates, so that a restarted run
        reproduces the uninterrupted one)."""
        self._base = posq.clone()

    def __call__(self, step, posq, force):
        import torch
        if self._base is None:
            self.set_base(posq)
        if self._noise is None:
            gen = torch.Generator(device=posq.device)
            gen.manual_seed(self.seed)
            self._noise = []
            for _ in range(self.fields):
                z = torch.zeros_like(posq)
                z[..., :3] = torch.clamp(torch.randn(posq.shape[:-1] + (3,), generator=gen, device=posq.device) * self.sigma,
                                         -2.5 * self.sigma, 2.5 * self.sigma)
                self._noise.append(z)
        torch.add(self._base, self._noise[step % self.fields], out=posq)


class ReplicaExchangeDriver:
    def __init__(self, system, schedule, jobname, outdir=".", temperature=300.0, steps_per_cycle=100, prune_every=5,
                 rebuild_every=40, skin=0.05, skin_outer=0.3, checkpoint_every=1, propagate=None, lambdas=None, seed=2022,
                 rank=0, world_size=1, device=0, group=None):
        import torch
        self.torch = torch
        self.s, self.jobname, self.outdir = system, jobname, outdir
        self.schedule = np.ascontiguousarray(schedule, np.float64)
        self.temperature = float(temperature)
        self.steps_per_cycle, self.prune_every, self.rebuild_every = int(steps_per_cycle), int(prune_every), int(rebuild_every)
        self.checkpoint_every = int(checkpoint_every)
        self.rank, self.world = int(rank), int(world_size)
        nrep = self.schedule.shape[0]
        self.lambdas = list(lambdas) if lambdas is not None else [float(row[1]) for row in self.schedule]
        self.rex = ReplicaExchange(self.schedule, nrep, rank=rank, world_size=world_size, temperature=temperature, seed=seed, group=group)
        self.mine = self.rex.mine
        self.R = len(self.mine)
        self.n = system["pos"].shape[0]
        os.makedirs(outdir, exist_ok=True)
        self.dev = torch.device("cuda", device)
        self.stream = torch.cuda.Stream(device=self.dev)
        self.be = ATMBackend(self.n, precision="mixed", num_replicas=max(self.R, 1), device=device)
        self.P = self.be.P
        self.be.set_displacements(system["displ"])
        self.be.set_box(system["box"])
        self.be.nb_setup(system["charge"], system["sigma"], system["epsilon"], system["cutoff"], system["ewald_alpha"], skin=skin,
                         skin_outer=skin_outer, exclusions=system.get("excl"), exception_pairs=system.get("exc14"),
                         exception_params=system.get("exc14_par"))
        posq = np.zeros((max(self.R, 1), self.P, 4), np.float32)
        posq[:, :self.n, :3] = system["pos"]
        posq[:, :self.n, 3] = system["charge"]
        self.posq = torch.from_numpy(posq).to(self.dev)
        self.force = torch.zeros((max(self.R, 1), 3 * self.P), dtype=torch.int64, device=self.dev)
        self.propagate = propagate if propagate is not None else JitterPropagator(seed=seed + 7919 * rank)
        if hasattr(self.propagate, "set_base"):
            self.propagate.set_base(self.posq)
        self.step_no = 0
        self.cycle = 0
        self._attached = False
        self._out = {}

    # ---- files ------------------------------------------------------------------------------------------------------
    def out_path(self, g):
        return os.path.join(self.outdir, f"{self.jobname}_r{g}.out")

    def chk_path(self, g):
        return os.path.join(self.outdir, f"{self.jobname}_r{g}-chk.xml")

    def hrex_path(self):
        return os.path.join(self.outdir, f"{self.jobname}-hrex.json")

    def _attach(self):
        for k, g in enumerate(self.mine):
            self.be.set_parameters(self.schedule[self.rex.replica_state[g]], replica=k)
        with self.torch.cuda.stream(self.stream):
            self.rex.attach_device(self.be, stream=self.stream)
        self._attached = True

    def restart(self):
        """Positions, step / cycle counters and the state permutation from the last checkpoint.  Returns False when
        there is none (a fresh start)."""
        if not os.path.exists(self.hrex_path()) or not all(os.path.exists(self.chk_path(g)) for g in self.mine):
            return False
        meta = json.load(open(self.hrex_path()))
        self.rex.load_state_dict(meta["hrex"])
        self.cycle, self.step_no = int(meta["cycle"]), int(meta["step"])
        posq = self.posq.cpu().numpy()
        for k, g in enumerate(self.mine):
            st = io.read_state_xml(self.chk_path(g))
            posq[k, :self.n, :3] = st["positions"]
        self.posq.copy_(self.torch.from_numpy(posq))
        self._attached = False
        return True

    def checkpoint(self):
        pos = self.posq.cpu().numpy()
        for k, g in enumerate(self.mine):
            row = self.schedule[self.rex.replica_state[g]]
            io.write_state_xml(self.chk_path(g), pos[k, :self.n, :3].astype(np.float64), self.s["box"], dict(zip(PARAM_NAMES, row)),
                               time=self.step_no * 1e-3)
        if self.rank == 0:
            with open(self.hrex_path(), "w") as fh:
                json.dump({"hrex": self.rex.state_dict(), "cycle": self.cycle, "step": self.step_no, "replicas": int(self.schedule.shape[0]),
                           "world_size": self.world}, fh)

    # ---- the loop ---------------------------------------------------------------------------------------------------
    def run(self, ncycles):
        """ncycles exchange cycles.  Returns the list of (cycle, replica, state, PE, u_sc) samples this rank wrote."""
        torch = self.torch
        if not self._attached:
            self._attach()
        samples = []
        for g in self.mine:
            if g not in self._out:
                self._out[g] = open(self.out_path(g), "a")
        for _ in range(ncycles):
            with torch.cuda.stream(self.stream):
                for _s in range(self.steps_per_cycle):
                    k = self.step_no
                    first = k == 0 or not self.be_has_lists
                    if first or k % self.rebuild_every == 0:
                        self.be.rebuild(self.posq, stream=self.stream)
                        self.be_has_lists = True
                    conc = not first and k % self.rebuild_every != 0 and k % self.prune_every == 0
                    self.be.step(self.posq, self.force, include_energy=True, graph=True, stream=self.stream, concurrent_prune=conc)
                    self.propagate(k, self.posq, self.force)
                    self.force.zero_()
                    self.step_no += 1
                # the sample of this cycle: one more evaluation at the cycle's final coordinates
                self.be.step(self.posq, self.force, include_energy=True, graph=True, stream=self.stream)
                self.force.zero_()
            en = self.be.get_energies(stream=self.stream)     # synchronises: once per cycle, to write the sample lines
            state = self.rex.sync_from_device(stream=self.stream) if self.cycle > 0 else self.rex.replica_state
            for k, g in enumerate(self.mine):
                row = self.schedule[state[g]]
                line = io.format_sample_line(self.temperature, self.lambdas[state[g]], row[0], row[1], row[2], row[3], row[4],
                                             float(en[k, 5]), float(en[k, 3]))
                self._out[g].write(line + "\n")
                self._out[g].flush()
                samples.append((self.cycle, g, int(state[g]), float(en[k, 5]), float(en[k, 3])))
            with torch.cuda.stream(self.stream):
                self.rex.exchange_device(stream=self.stream)
            self.cycle += 1
            if self.checkpoint_every > 0 and self.cycle % self.checkpoint_every == 0:
                self.rex.sync_from_device(stream=self.stream)
                self.checkpoint()
        return samples

    be_has_lists = False

    def close(self):
        for f in self._out.values():
            f.close()
        self._out = {}
        self.be.close()
