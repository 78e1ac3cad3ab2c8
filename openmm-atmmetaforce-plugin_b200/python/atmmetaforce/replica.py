"""Lambda-window replica layer: partition of the replicas over the ranks of one node and the Hamiltonian
replica-exchange step.  The reference has no counterpart: one OpenMM Context holds one lambda
state and the user moves lambdas with context.setParameter(...).

Data flow per exchange cycle: every rank contributes the (U1, U2) of its resident replicas (16 B per replica) to one
all-gather (NCCL over NVLink on GPUs, gloo in the CPU tests); every rank then runs the SAME deterministic Metropolis
sweep (counter-based RNG keyed by (seed, cycle), C ABI atm_hrex_sweep) and swaps lambda-STATE indices -- coordinates
never move between GPUs.
"""
import numpy as np

from .backend import hrex_sweep
from .synthetic import partition_replicas

KB = 0.0083144626  # kJ/(mol K)


class ReplicaExchange:
    def __init__(self, schedule, num_replicas, rank=0, world_size=1, temperature=300.0, seed=2022, group=None):
        self.schedule = np.ascontiguousarray(schedule, np.float64)
        self.num_replicas = int(num_replicas)
        self.rank, self.world = int(rank), int(world_size)
        self.parts = partition_replicas(self.num_replicas, self.world)
        self.mine = self.parts[self.rank]
        self.max_per_rank = max(len(p) for p in self.parts)
        self.beta = 1.0 / (KB * temperature)
        self.seed = int(seed)
        self.group = group
        self.cycle = 0
        self.accepted = 0
        self.attempted_cycles = 0
        # replica g starts in state g (mod number of states)
        self.replica_state = (np.arange(self.num_replicas) % self.schedule.shape[0]).astype(np.int32)
        self._slot = {}
        for r, lst in enumerate(self.parts):
            for k, g in enumerate(lst):
                self._slot[g] = r * self.max_per_rank + k

    def local_parameters(self):
        """[len(mine)][9] parameter rows of the replicas resident on this rank."""
        return self.schedule[self.replica_state[self.mine]]

    def gather(self, local_u12):
        """all-gather of the per-replica (U1, U2): local_u12 is a [len(mine), 2] float64 torch tensor on the device the
        process group works with (or a numpy array when world_size == 1).  Returns a [num_replicas, 2] numpy array."""
        if self.world == 1:
            arr = local_u12.detach().cpu().numpy() if hasattr(local_u12, "detach") else np.asarray(local_u12)
            return np.ascontiguousarray(arr[:len(self.mine)], np.float64)
        import torch
        import torch.distributed as dist
        send = torch.zeros((self.max_per_rank, 2), dtype=torch.float64, device=local_u12.device)
        send[:len(self.mine)] = local_u12[:len(self.mine)]
        recv = torch.zeros((self.world * self.max_per_rank, 2), dtype=torch.float64, device=local_u12.device)
        dist.all_gather_into_tensor(recv, send, group=self.group)
        allu = recv.cpu().numpy()
        return np.stack([allu[self._slot[g]] for g in range(self.num_replicas)])

    def exchange(self, local_u12):
        """One cycle.  Returns the list of (local replica index, new parameter row) whose state changed."""
        u12 = self.gather(local_u12)
        if not np.all(np.isfinite(u12)):
            raise FloatingPointError("non-finite replica energy in the exchange step")
        self.cycle += 1
        new_state, acc = hrex_sweep(self.schedule, u12, self.replica_state, self.beta, self.seed, self.cycle)
        changed = [(k, self.schedule[new_state[g]]) for k, g in enumerate(self.mine) if new_state[g] != self.replica_state[g]]
        self.replica_state[:] = new_state
        self.accepted += acc
        self.attempted_cycles += 1
        return changed

    # ---- the same cycle without a host round trip (GPU back-end only)
    def attach_device(self, backend, stream=None, collective="library"):
        """Move the exchange bookkeeping onto the device of `backend` (an ATMBackend holding this rank's replicas in the
        order of self.mine).  After this, exchange_device() runs a whole cycle asynchronously on the stream:
        pack kernel -> all-gather (NCCL) -> sweep kernel that rewrites the local parameter rows in place.
        collective = "library": the all-gather is issued inside libatm_b200 on a communicator of its own
        (atm_re_comm_create / atm_hrex_device_cycle; the NCCL unique id travels through torch.distributed once);
        "torch": torch.distributed.all_gather_into_tensor between the library's pack and sweep kernels."""
        import torch
        self._be = backend
        self._collective = collective
        self._comm = None
        if collective == "library":
            from .backend import ReplicaComm

            def share(raw):
                import torch.distributed as dist
                box = [raw]
                dist.broadcast_object_list(box, src=0, group=self.group)
                return box[0]
            try:
                self._comm = ReplicaComm(self.rank, self.world, torch.cuda.current_device(), share if self.world > 1 else None)
                ok = 1
            except Exception as exc:  # no usable libnccl for the library (e.g. a torch build without NCCL): say so, use torch's
                import sys
                print(f"atmmetaforce: library NCCL communicator unavailable ({exc}); using torch.distributed for the all-gather", file=sys.stderr)
                ok = 0
            if self.world > 1:   # every rank must take the same path
                import torch.distributed as dist
                flag = torch.tensor([ok], device=torch.device("cuda", torch.cuda.current_device()))
                dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
                ok = int(flag.item())
            if not ok:
                self._collective = "torch"
                self._comm = None
        rows = self.world * self.max_per_rank
        gather_slot = np.array([self._slot[g] for g in range(self.num_replicas)], np.int32)
        backend.hrex_setup(self.schedule, self.replica_state, self.mine, gather_slot, rows, self.beta, self.seed, stream=stream)
        dev = torch.device("cuda", torch.cuda.current_device())
        self._send = torch.zeros((self.max_per_rank, 2), dtype=torch.float64, device=dev)
        self._recv = torch.zeros((rows, 2), dtype=torch.float64, device=dev) if self.world > 1 else self._send

    def exchange_device(self, stream=None):
        """One cycle, fully asynchronous (no .cpu(), no synchronisation).  Bookkeeping: sync_from_device()."""
        self.cycle += 1
        self.attempted_cycles += 1
        if self._collective == "library":
            self._be.hrex_cycle(self._comm, self.cycle, stream=stream)
            return
        import torch
        self._be.hrex_pack(self._send, stream=stream)
        if self.world > 1:
            import torch.distributed as dist
            # the collective is ordered on torch's CURRENT stream: make that the stream of the pack and sweep kernels
            with torch.cuda.stream(stream) if stream is not None else torch.cuda.stream(torch.cuda.current_stream()):
                dist.all_gather_into_tensor(self._recv, self._send, group=self.group)
        self._be.hrex_exchange(self._recv, self.cycle, stream=stream)

    def sync_from_device(self, stream=None):
        """Pull the state permutation and the acceptance count back (synchronises the stream)."""
        rs, acc, cycles, err = self._be.hrex_state(self.num_replicas, stream=stream)
        if err:
            raise FloatingPointError("non-finite replica energy in an exchange step (that cycle was skipped)")
        self.replica_state[:] = rs
        self.accepted = acc
        return rs

    def state_dict(self):
        """Checkpoint of the exchange bookkeeping (state permutation, RNG counter).  With the bookkeeping on the device
        the permutation is pulled back first (synchronises)."""
        if getattr(self, "_be", None) is not None:
            self.sync_from_device()
        return {"replica_state": self.replica_state.tolist(), "cycle": self.cycle, "seed": self.seed,
                "accepted": self.accepted}

    def load_state_dict(self, d):
        self.replica_state[:] = np.asarray(d["replica_state"], np.int32)
        self.cycle, self.seed, self.accepted = int(d["cycle"]), int(d["seed"]), int(d.get("accepted", 0))
        if getattr(self, "_be", None) is not None:
            # the device holds its own copy of the permutation and the parameter rows: upload the restored ones
            gather_slot = np.array([self._slot[g] for g in range(self.num_replicas)], np.int32)
            self._be.hrex_setup(self.schedule, self.replica_state, self.mine, gather_slot, self.world * self.max_per_rank,
                                self.beta, self.seed)
            for k, row in enumerate(self.local_parameters()):
                self._be.set_parameters(row, replica=k)
