#!/usr/bin/env python
"""bench.py -- ATM Meta-Force hot path: aggregate replica-ns/day on N B200 GPUs of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W]            (N > 1: launched by torchrun, one rank per GPU)
    python bench.py --impl reference [--steps K] [--warmup W]      (the CPU arm: the oracle port on the host cores)

Workload (BASELINE.json configs[2]): synthetic ~23k-atom solvated blob + 40-atom ligand (TIP3P-like water, Ewald real
space, 0.9 nm cutoff), 22-state ABFE softplus lambda schedule = 22 replicas, block-cyclically partitioned over the
ranks (no data-path collective; an NCCL all-gather of the per-replica (U1,U2) every --exchange-every steps feeds the
Hamiltonian replica-exchange sweep).  One "step" = one pass of the hot path (copy-state/pack -> two-state direct-space
nonbonded -> device scalar stage -> merge) for every replica resident on the rank, plus the pair-list maintenance on
its declared cadence (prune every --prune-every steps, full rebuild every --rebuild-every steps).

metric = aggregate replica-ns/day = 22 replicas * dt / (max-over-ranks time per step); dt = 1 fs as in the reference's
example scripts (example/abfe/abfe.py:94).  OpenMM's PME reciprocal space, bonded terms and integrator are NOT part of
the measured path (they stay in OpenMM, which is not installable here) -- see DESIGN.md "Measurement".
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "openmm-atmmetaforce-plugin_b200", "python"))

import numpy as np  # noqa: E402

DT_FS = 1.0
NUM_REPLICAS = 22
FP32_PEAK_TFLOPS = None  # derived from the device at run time: SMs * 128 lanes * 2 flop * max SM clock
FLOP_PER_PAIR = 60.0
JITTER_NM = 0.005        # per-step seeded Gaussian position noise around the base coordinates, clipped at 2.5 sigma


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="config3", choices=["config3", "config4", "rbfe", "abfe"])
    ap.add_argument("--replicas", type=int, default=NUM_REPLICAS)
    ap.add_argument("--dt-fs", type=float, default=1.0, help="MD time step the ns/day conversion uses (reference examples: "
                    "1 fs; 4 fs = HMR practice).  The default pair-list cadences scale with it: the same physical time "
                    "between prunes (5 fs, inner skin 0.05 nm) and rebuilds (40 fs, outer skin 0.3 nm)")
    ap.add_argument("--prune-every", type=int, default=0, help="steps between prunes (0 = 5 fs / dt)")
    ap.add_argument("--rebuild-every", type=int, default=0, help="steps between rebuilds (0 = 40 fs / dt)")
    ap.add_argument("--exchange-every", type=int, default=100)
    ap.add_argument("--skin", type=float, default=0.05, help="inner pair-list skin in nm; the default prune cadence (5 fs) keeps the "
                    "same margin per unit of time as 0.1 nm / 10 fs (measured: 0.491 vs 0.507 ms per step at 22 replicas)")
    ap.add_argument("--skin-outer", type=float, default=0.3)
    ap.add_argument("--host-exchange", action="store_true", help="run the replica-exchange sweep on the host (D2H copy + "
                    "synchronisation per cycle) instead of the on-device cycle")
    ap.add_argument("--prune-mode", default="concurrent", choices=["concurrent", "before"], help="concurrent: the prune of step "
                    "k's coordinates runs on the back-end's side stream during step k and serves steps k+1... (atm_step_io."
                    "concurrent_prune); before: atm_nb_prune on the launching stream before the step (round 1)")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-flush", action="store_true")
    ap.add_argument("--flush-mode", default="write", choices=["write", "write+read"], help="L2 flush between steps (outside "
                    "the event pairs).  write: a 256 MiB memset.  write+read: the memset followed by a read of another 256 MiB "
                    "buffer, so that the cache holds CLEAN lines of unrelated data (measured: no difference -- 0.4260 / 0.4272 ms "
                    "per step at 22 replicas, 0.0770 / 0.0757 at 3: the write-back of the memset's dirty lines is not charged "
                    "to the timed kernels)")
    ap.add_argument("--cpu-steps", type=int, default=200, help="upper bound of the CPU-baseline sample (also capped at ~15 s)")
    ap.add_argument("--cpu-budget-s", type=float, default=150.0, help="--impl reference: time budget of the whole run; a step "
                    "covers fewer replicas when K steps of all of them would not fit")
    ap.add_argument("--skip-two-separate", action="store_true")
    ap.add_argument("--skip-tier1", action="store_true", help="skip the 8M-atom HBM roofline probe of copy-state / hybrid-force")
    ap.add_argument("--skip-e2e", action="store_true", help="profiling runs only: leave the host-buffer leg out (its chunk handles "
                    "add their own rebuilds and small launches to an ncu launch list)")
    ap.add_argument("--e2e-chunks", type=int, default=6)
    ap.add_argument("--e2e-split", default="auto",
                    help="replicas per chunk of the e2e leg: 'equal' (--e2e-chunks equal chunks), 'graded' (seven chunks 1:2:3:6:6:3:1, "
                         "small first and last chunks because their upload / download is not hidden), an explicit list such as "
                         "'1,2,4,8,5,2', or 'auto' (default): graded from 16 replicas per rank, equal below -- measured "
                         "3 %% faster than equal at 22 replicas and slower at 11 / 6 / 3 (profiles/r2u_e2e_split_sweep.log)")
    ap.add_argument("--e2e-posq", default="f3", choices=["f3", "f4"], help="what the e2e leg uploads: packed float3 coordinates "
                    "(12 B per atom, ATM_POSQ_F3) or OpenMM's float4 posq (16 B per atom)")
    ap.add_argument("--e2e-force", default="f32", choices=["f32", "i64"], help="what the e2e leg reads back: float32 forces "
                    "(12 B per atom) or the 2^32 fixed-point long force buffer (24 B per atom, round 1)")
    ap.add_argument("--pme", action="store_true", help="also evaluate the two-state PME reciprocal space inside the step "
                    "(SURVEY 8f row 1; NOT part of the headline workload, which is the direct-space path)")
    args = ap.parse_args()
    global DT_FS
    DT_FS = args.dt_fs
    if args.prune_every <= 0:
        args.prune_every = max(1, int(round(5.0 / DT_FS)))
    if args.rebuild_every <= 0:
        args.rebuild_every = max(args.prune_every, int(round(40.0 / DT_FS)))
    return args


def load_workload(name):
    from atmmetaforce import synthetic
    if name == "config3":
        s = synthetic.config3()
        label = "synthetic 23k-atom solvated blob + 40-atom ligand, 22-state ABFE softplus schedule (BASELINE configs[2])"
    elif name == "config4":
        s = synthetic.config4()
        label = "synthetic 100k-atom box, two 40-atom ligands (RBFE), 22 replicas (BASELINE configs[3])"
    else:
        g = dict(np.load(os.path.join(ROOT, "tests", "golden",
                                      "temoa_g1_abfe.npz" if name == "abfe" else "temoa_g1_g4_rbfe.npz")))
        g["cutoff"] = 1.0
        g["ewald_alpha"] = synthetic.ewald_alpha(1.0)
        s = g
        label = ("TEMOA-G1 ABFE fixture (BASELINE configs[0])" if name == "abfe"
                 else "TEMOA-G1/G4 RBFE fixture, two displacement groups (BASELINE configs[1])")
    return s, synthetic.atm_schedule_22(), label


def replica_positions(s, replica):
    """Replica k = base coordinates + seeded 0.002 nm noise (the same for any number of ranks)."""
    rng = np.random.default_rng(1000 + replica)
    pos = s["pos"].copy()
    if replica > 0:
        pos += rng.normal(0.0, 0.002, pos.shape)
    return pos


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def _oracle_system(s):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_py as O
    # every core this process may run on (torchrun exports OMP_NUM_THREADS=1 to its workers)
    O.set_num_threads(len(os.sched_getaffinity(0)))
    S = O.System(s["charge"], s["sigma"], s["epsilon"], s["box"], s["cutoff"], s["ewald_alpha"], s["excl"],
                 s["exc14"], s["exc14_par"])
    return O, S


def cpu_oracle_rate(s, sched, steps, budget_s=15.0):
    """cpu_baseline leg: the CPU restatement of the same hot path (oracle port, OpenMP, all host threads) on ONE
    replica; returns (replica-ns/day, seconds per replica-step, threads, steps done)."""
    O, S = _oracle_system(s)
    pos = replica_positions(s, 0)
    S.step(sched[0], pos, s["displ"])  # warm-up (page-in, OpenMP pool)
    t0 = time.perf_counter()
    done = 0
    for k in range(steps):
        S.step(sched[k % len(sched)], pos, s["displ"])
        done += 1
        if time.perf_counter() - t0 > budget_s:   # bounded sample: stop after ~budget_s seconds of CPU work
            break
    dt = (time.perf_counter() - t0) / done
    return DT_FS * 1e-6 * 86400.0 / dt, dt, O.num_threads(), done


def tier1_hbm_probe(torch, atm, dev, flush, atoms=8_000_000):
    """HBM roofline of the two Tier-1 kernels (north_star: copy + merge >= 70 % of peak), measured live on a working
    set far beyond L2: CopyState in mixed precision (112 B/atom) and HybridForce (96 B/atom), median of 10 launches,
    CUDA events on the launching stream, L2 flushed between launches.  Peak = MEASURED_PEAKS.json hbm_gbs."""
    try:
        peak, src = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"], "measured"
    except Exception:
        peak, src = 6650.0, "fallback"
    n = atoms
    P = 32 * ((n + 31) // 32)
    d = np.zeros((n, 3)); d[:50] = [2.2, 2.2, 2.2]
    be = atm.ATMBackend(n, padded_num_particles=P, precision="mixed", device=dev.index)
    be.set_displacements(d)
    posq = torch.rand((P, 4), device=dev)
    c = torch.zeros_like(posq)
    p1, p2, c1, c2 = (torch.empty_like(posq) for _ in range(4))
    f0 = torch.zeros(3 * P, dtype=torch.int64, device=dev)
    f1 = torch.randint(-2**40, 2**40, (3 * P,), dtype=torch.int64, device=dev)
    f2 = torch.randint(-2**40, 2**40, (3 * P,), dtype=torch.int64, device=dev)
    out = {"atoms": n, "peak_GBps": peak, "peak_source": src, "l2": "flushed between launches"}
    for name, bpa, fn in (("copy_state_mixed", 112, lambda: be.copy_state(posq, p1, p2, c, c1, c2)),
                          ("hybrid_force", 96, lambda: be.hybrid_force(f0, f1, f2, 0.37))):
        ts = []
        for it in range(13):
            if flush is not None:
                flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record()
            torch.cuda.synchronize()
            if it >= 3:
                ts.append(a.elapsed_time(b))
        ms = float(np.median(ts))
        out[name] = {"bytes": bpa * n, "ms": ms, "GBps": bpa * n / (ms * 1e-3) / 1e9, "frac": bpa * n / (ms * 1e-3) / 1e9 / peak}
    be.close()
    return out


def e2e_split(spec, R, nchunks):
    """Replicas per chunk of the e2e leg (see --e2e-split)."""
    if spec not in ("auto", "graded", "equal"):
        sizes = [int(x) for x in spec.split(",")]
        if sum(sizes) != R or min(sizes) < 1:
            raise SystemExit(f"--e2e-split {spec}: the chunk sizes must be positive and add up to the {R} replicas of this rank")
        return sizes
    k = max(1, min(nchunks, R))
    if spec == "equal" or (spec == "auto" and R < 16) or R < 7:
        b = [round(i * R / k) for i in range(k + 1)]
        return [b[i + 1] - b[i] for i in range(k)]
    # graded: seven chunks in the proportions 1 : 2 : 3 : 6 : 6 : 3 : 1 -- the first upload and the last download are not
    # hidden behind any compute, so those chunks are small; the central ones are large for per-launch efficiency
    w = [1, 2, 3, 6, 6, 3, 1]
    sizes = [max(1, (x * R) // 22) for x in w]
    rest = R - sum(sizes)
    i = 0
    while rest != 0:                       # the remainder goes to (or comes from) the two central chunks
        j = 3 + (i & 1)
        step = 1 if rest > 0 else -1
        if sizes[j] + step >= 1:
            sizes[j] += step
            rest -= step
        i += 1
    return sizes


def workload_config(args, label, s, replicas_per_rank, exchange, use_graph, pme_grid, flush_note):
    """The `config` object of the JSON line; both arms print the same keys."""
    return {"workload": label, "replicas": args.replicas, "replicas_per_rank": replicas_per_rank, "atoms": int(s["pos"].shape[0]),
            "dt_fs": DT_FS, "cutoff_nm": s["cutoff"], "skin_nm": args.skin, "skin_outer_nm": args.skin_outer,
            "prune_every": args.prune_every, "rebuild_every": args.rebuild_every, "exchange_every": args.exchange_every,
            "prune_mode": args.prune_mode, "exchange": exchange, "cuda_graph": use_graph, "pme_grid": pme_grid, "l2": flush_note,
            "jitter_nm": JITTER_NM}


def run_reference(args):
    """The reference arm: the reference's own CPU path cannot be built here (every translation unit needs OpenMM,
    which is absent from the image), so this times the oracle port -- a C/OpenMP restatement of the same path -- with
    all host threads.  One step = every replica of the schedule evaluated once (two full direct-space evaluations +
    scalar stage + merge each), one after another: the same whole-job step the GPU arm times.  When K steps of all
    replicas would not fit the time budget, a step covers the first `sample` replicas only and the line says so."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    s, sched, label = load_workload(args.workload)
    O, S = _oracle_system(s)
    R = args.replicas
    pos = [replica_positions(s, g) for g in range(R)]
    t0 = time.perf_counter()
    S.step(sched[0], pos[0], s["displ"])          # page-in + OpenMP pool + a first estimate of the cost
    S.step(sched[0], pos[0], s["displ"])
    est = (time.perf_counter() - t0) / 2
    W, K = max(0, args.warmup), max(1, args.steps)
    budget = args.cpu_budget_s
    sample = R
    if est * R * (K + W) > budget:                # bounded sample: fewer replicas per step, never fewer steps
        sample = max(1, min(R, int(budget / (est * (K + W)))))
    rng = np.random.default_rng(2022)

    def one_step(k):
        for g in range(sample):
            x = pos[g] + np.clip(rng.normal(0.0, JITTER_NM, pos[g].shape), -2.5 * JITTER_NM, 2.5 * JITTER_NM)
            S.step(sched[g % len(sched)], x, s["displ"])

    for k in range(W):
        one_step(k)
    t0 = time.perf_counter()
    for k in range(K):
        one_step(k)
    sec = (time.perf_counter() - t0) / K          # seconds per step of `sample` replicas
    rate = sample * DT_FS * 1e-6 * 86400.0 / sec  # replicas are evaluated one after another: aggregate == per-replica rate
    threads = O.num_threads()
    line = {
        "impl": "reference", "metric": "aggregate replica-ns/day (ATM hot path)", "value": rate, "unit": "replica-ns/day",
        "n_gpus": args.gpus, "steps": K, "warmup": W, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, label, s, None, None, None, None, None),
        "replicas_per_step_timed": sample,
        "ms_per_step_all_replicas": sec * 1e3 * R / sample,
        "cpu_baseline": {"value": rate, "unit": "replica-ns/day", "cores": threads, "kind": "port",
                         "sample": f"{K} steps (+{W} warm-up) of {sample} of the {R} replicas, evaluated one after another with "
                                   f"{threads} OpenMP threads (two full direct-space evaluations + scalar stage + merge per replica-step)"},
        "e2e": {"value": rate, "unit": "replica-ns/day", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "oracle/_ref (the reference compiled in place) is not buildable: OpenMM is absent; kind=port",
    }
    emit(line)


class _DevView:
    """CUDA-array-interface view of library-owned device memory (for torch.as_tensor)."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": shape, "typestr": typestr, "data": (ptr, False), "version": 2}


def run_b200(args):
    import torch
    import torch.distributed as dist
    import atmmetaforce as atm
    from atmmetaforce import synthetic, _capi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 arm has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    s, sched, label = load_workload(args.workload)
    n = s["pos"].shape[0]
    total_replicas = args.replicas
    rex = atm.ReplicaExchange(sched, total_replicas, rank=rank, world_size=world, temperature=300.0, seed=2022)
    mine = rex.mine
    R = len(mine)
    max_per_rank = rex.max_per_rank
    replica_state = rex.replica_state

    be = atm.ATMBackend(n, precision="mixed", num_replicas=max(R, 1), device=local_rank)
    P = be.P
    be.set_displacements(s["displ"])
    be.set_box(s["box"])
    for k, g in enumerate(mine):
        be.set_parameters(sched[replica_state[g]], replica=k)
    be.nb_setup(s["charge"], s["sigma"], s["epsilon"], s["cutoff"], s["ewald_alpha"], skin=args.skin,
                skin_outer=args.skin_outer, exclusions=s["excl"], exception_pairs=s["exc14"],
                exception_params=s["exc14_par"])
    pme_grid = None
    if args.pme:
        pme_grid = synthetic.pme_grid(s["box"], s["ewald_alpha"])
        be.pme_setup(pme_grid)
    posq_h = torch.zeros((max(R, 1), P, 4), dtype=torch.float32).pin_memory()
    corr_h = torch.zeros((max(R, 1), P, 4), dtype=torch.float32).pin_memory()
    for k, g in enumerate(mine):
        p64 = replica_positions(s, g)
        p32 = p64.astype(np.float32)
        posq_h[k, :n, :3] = torch.from_numpy(p32)
        posq_h[k, :n, 3] = torch.from_numpy(s["charge"].astype(np.float32))
        corr_h[k, :n, :3] = torch.from_numpy((p64 - p32).astype(np.float32))
    posq = posq_h.to(dev)
    corr = corr_h.to(dev)
    force = torch.zeros((max(R, 1), 3 * P), dtype=torch.int64, device=dev)
    force_h = torch.zeros((max(R, 1), 3 * P), dtype=torch.float32 if args.e2e_force == "f32" else torch.int64).pin_memory()
    stream = torch.cuda.Stream(device=dev)
    use_graph = not args.no_graph
    class _Flush:
        """L2 flush: write a buffer twice the size of L2, then (write+read) read a second one back."""

        def __init__(self):
            self.w = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
            self.r = torch.zeros(64 * 1024 * 1024, dtype=torch.int32, device=dev) if args.flush_mode == "write+read" else None
            self.sink = torch.zeros(1, dtype=torch.int64, device=dev)

        def zero_(self):
            self.w.zero_()
            if self.r is not None:
                torch.sum(self.r, dim=0, keepdim=True, out=self.sink)

    flush = None if args.no_flush else _Flush()
    device_exchange = not args.host_exchange and total_replicas >= world
    if device_exchange:
        with torch.cuda.stream(stream):
            rex.attach_device(be, stream=stream)

    def exchange(cycle):
        """all-gather (U1,U2) of every replica, identical Metropolis sweep on every rank, swap lambda states."""
        if device_exchange:   # pack kernel -> NCCL all-gather -> sweep kernel, no host synchronisation
            rex.exchange_device(stream=stream)
            return
        en = torch.as_tensor(_DevView(be.energies_device_ptr(), (max(R, 1), _capi.NUM_ENERGY_SLOTS), "<f8"), device=dev)
        for k, row in rex.exchange(en[:, 0:2].contiguous()):
            be.set_parameters(row, replica=k)

    # ---- per-step position jitter: x = base + clipped N(0, 0.005 nm), a seeded set of NJ noise
    #      fields cycled through, applied on the device OUTSIDE the event pairs; the prune / rebuild therefore see
    #      different coordinates every time (round 1 re-evaluated identical coordinates)
    NJ = 8
    gen = torch.Generator(device=dev)
    gen.manual_seed(2022 + rank)
    base = posq.clone()
    noise = []
    for j in range(NJ):
        z = torch.zeros_like(posq)
        z[:, :n, :3] = torch.clamp(torch.randn((max(R, 1), n, 3), generator=gen, device=dev) * JITTER_NM, -2.5 * JITTER_NM, 2.5 * JITTER_NM)
        noise.append(z)

    # steady-state frequencies of the step kinds over one period of the declared cadences
    period = int(np.lcm(args.prune_every, args.rebuild_every))
    n_reb = sum(1 for k in range(period) if k % args.rebuild_every == 0)
    n_pru = sum(1 for k in range(period) if k % args.rebuild_every != 0 and k % args.prune_every == 0)
    freq = {"rebuild": n_reb / period, "prune": n_pru / period, "exchange": 1.0 / args.exchange_every if total_replicas > 1 else 0.0}

    step_no = [0]

    def one_step(ev=None, kind=None, do_ex=None):
        """One pass of the hot path for every resident replica.  ev = 4 events: start | after pair-list maintenance |
        after the step | after the exchange cycle."""
        with torch.cuda.stream(stream):
            k = step_no[0]
            torch.add(base, noise[k % NJ], out=posq)
            if flush is not None:
                flush.zero_()
            if kind is None:
                kind = "rebuild" if k % args.rebuild_every == 0 else ("prune" if k % args.prune_every == 0 else "plain")
            if do_ex is None:
                do_ex = (k + 1) % args.exchange_every == 0
            if ev is not None:
                ev[0].record(stream)
            conc = kind == "prune" and args.prune_mode == "concurrent"
            if kind == "rebuild":
                be.rebuild(posq, stream=stream)
            elif kind == "prune" and not conc:
                be.prune(posq, stream=stream)
            if ev is not None:
                ev[1].record(stream)
            be.step(posq, force, posq_corr=corr, include_energy=True, graph=use_graph, stream=stream, concurrent_prune=conc)
            if ev is not None:
                ev[2].record(stream)
            if do_ex:
                exchange((k + 1) // args.exchange_every)
            if ev is not None:
                ev[3].record(stream)
            step_no[0] += 1
        return kind, do_ex

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def new_events():
        return [torch.cuda.Event(enable_timing=True) for _ in range(4)]

    # ---- warm-up (>= 3 steps), then one statistics step (pair counts; not timed)
    W = max(3, args.warmup)
    if R > 0:
        for _ in range(W):
            one_step()
        # priming (untimed, on top of the W warm-up steps): every cached CUDA graph the timed region can replay is
        # captured here -- the step on either copy of the pruned list, with and without a concurrent prune, the
        # asynchronous rebuild -- so that no capture / instantiation lands inside an event pair of a short window
        for kind_p in ("prune", "plain", "prune", "plain", "rebuild", "plain", "prune", "plain"):
            one_step(None, kind=kind_p, do_ex=False)
    # one untimed exchange cycle: the first NCCL collective builds the communicator (tens of ms), which is set-up
    # cost, not steady-state step time
    with torch.cuda.stream(stream):
        exchange(0)
    if R > 0:
        with torch.cuda.stream(stream):
            be.step(posq, force, posq_corr=corr, include_energy=True, collect_stats=True, stream=stream)
        stats_en = be.get_energies(stream=stream)
    else:
        stats_en = np.zeros((1, _capi.NUM_ENERGY_SLOTS))
    nb_stats = be.nb_stats() if R > 0 else {}

    # ---- timed region: exactly K steps, per-step CUDA events on the launching stream (jitter and L2 flush stay outside
    #      the event pairs), barrier + synchronize on both sides, max over ranks
    K = args.steps
    records = []
    launches0 = be.launch_count()
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    t_wall0 = time.perf_counter()
    for k in range(K):
        ev = new_events()
        if R > 0:
            kind, ex = one_step(ev)
        else:  # a rank without replicas still takes part in the exchange collective
            kind, ex = "plain", (step_no[0] + 1) % args.exchange_every == 0
            with torch.cuda.stream(stream):
                for e in ev[:3]:
                    e.record(stream)
                if ex:
                    exchange(0)
                ev[3].record(stream)
            step_no[0] += 1
        records.append((kind, ex, ev))
    barrier()
    wall = time.perf_counter() - t_wall0
    clocks = sampler.stop() if rank == 0 else None
    launches = be.launch_count() - launches0
    window_ms = sum(ev[0].elapsed_time(ev[3]) for _, _, ev in records)

    # ---- the K-step window holds whatever maintenance / exchange calls fall on its step numbers (a 20-step window:
    #      4 prunes, at most one rebuild, no exchange), so the reported step time is the STEADY-STATE one:
    #          plain step + sum over kinds of (mean cost of that call) * (its frequency on the declared cadence),
    #      every term measured in this run with the same events; kinds the window saw fewer than 3 times are sampled
    #      right after it (outside the window, same stream, same jitter/flush discipline)
    extra = []
    barrier()   # rank 0 has just spent 0.15 s stopping the clock sampler: line the ranks up again before the collectives below
    for kind_x in ("prune", "rebuild"):
        have = sum(1 for kd, _, _ in records if kd == kind_x)
        for _ in range(max(0, 3 - have) if freq[kind_x] > 0 else 0):
            ev = new_events()
            if R > 0:
                one_step(ev, kind=kind_x, do_ex=False)
                one_step(None, kind="plain", do_ex=False)
                extra.append((kind_x, False, ev))
    have_ex = sum(1 for _, ex, _ in records if ex)
    for _ in range(max(0, 3 - have_ex) if freq["exchange"] > 0 else 0):
        ev = new_events()
        if R > 0:
            one_step(ev, kind="plain", do_ex=True)
        else:
            with torch.cuda.stream(stream):
                for e in ev[:3]:
                    e.record(stream)
                exchange(0)
                ev[3].record(stream)
        extra.append(("plain", True, ev))
    barrier()
    if device_exchange:
        rex.sync_from_device(stream=stream)   # bookkeeping only (raises on a non-finite energy), outside the timed region
    allrec = records + extra

    def mean_ms(sel, a, b):
        """Typical cost of one component: the mean over the plain steps' many samples; the MEDIAN for the few samples of
        a maintenance / exchange call, so that a one-off event (a capacity growth that reallocates the lists, a graph
        re-capture, NCCL's lazy connection set-up) does not pass for steady state."""
        v = [ev[a].elapsed_time(ev[b]) for kd, ex, ev in allrec if sel(kd, ex)]
        if not v:
            return 0.0, 0
        return (float(np.median(v)) if len(v) < 50 else sum(v) / len(v)), len(v)

    concurrent = args.prune_mode == "concurrent"
    # the step itself: every step, or -- with concurrent prunes, whose cost sits INSIDE the step interval -- the steps
    # without a prune; a prune step is then charged the difference
    t_plain, n_plain = mean_ms((lambda kd, ex: kd != "prune") if concurrent else (lambda kd, ex: True), 1, 2) if R > 0 else (0.0, 0)
    comp = {"step": {"ms": t_plain, "samples": n_plain, "per_step": 1.0}}
    steady = t_plain
    for kind_x in ("prune", "rebuild"):
        if kind_x == "prune" and concurrent:
            t, c = mean_ms(lambda kd, ex: kd == "prune", 0, 2)
            t = max(0.0, t - t_plain)
            comp[kind_x] = {"ms": t, "samples": c, "per_step": freq[kind_x], "note": "concurrent with the step: (prune step) - (plain step)"}
        else:
            t, c = mean_ms(lambda kd, ex, kx=kind_x: kd == kx, 0, 1)
            comp[kind_x] = {"ms": t, "samples": c, "per_step": freq[kind_x]}
        steady += t * freq[kind_x]
    t, c = mean_ms(lambda kd, ex: ex, 2, 3)
    comp["exchange"] = {"ms": t, "samples": c, "per_step": freq["exchange"]}
    steady += t * freq["exchange"]
    if os.environ.get("ATM_BENCH_DEBUG"):
        print(f"[bench rank {rank}] steady {steady:.4f} window {window_ms / K:.4f} " +
              " ".join(f"{k}={v['ms']:.4f}x{v['samples']}" for k, v in comp.items()), file=sys.stderr, flush=True)
    tt = torch.tensor([steady, window_ms / K], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms_per_step = float(tt[0].item())
    window_ms_per_step = float(tt[1].item())
    value = total_replicas * DT_FS * 1e-6 * 86400.0 / (ms_per_step * 1e-3)

    # ---- roofline of the dominant kernel (nb2): CUDA events around the launch, 20 more steps right after the timed region
    nb2_ms = None
    if R > 0:
        be.profile_enable(True)
        with torch.cuda.stream(stream):
            for _ in range(20):
                if flush is not None:
                    flush.zero_()
                be.step(posq, force, posq_corr=corr, include_energy=True, graph=False, stream=stream)
            torch.add(base, noise[0], out=posq)
        tot, cnt = be.profile_read()
        be.profile_enable(False)
        nb2_ms = tot / max(cnt, 1)

    # ---- north_star comparison: the fused two-state launch against TWO separate single-state evaluations (what two
    #      inner contexts do), measured with this library's own kernel on zero-displacement handles at x and at x+d
    two_sep = None
    if R > 0 and world == 1 and not args.skip_two_separate:
        zero = np.zeros_like(s["displ"])
        d32 = torch.from_numpy(s["displ"].astype(np.float32)).to(dev)
        posq2 = posq.clone()
        posq2[:, :n, :3] += d32  # the CopyState float add
        singles = []
        for pq in (posq, posq2):
            b1 = atm.ATMBackend(n, precision="mixed", num_replicas=R, device=local_rank)
            b1.set_displacements(zero)
            b1.set_box(s["box"])
            b1.set_parameters(sched[0])
            b1.nb_setup(s["charge"], s["sigma"], s["epsilon"], s["cutoff"], s["ewald_alpha"], skin=args.skin,
                        skin_outer=args.skin_outer, exclusions=s["excl"], exception_pairs=s["exc14"],
                        exception_params=s["exc14_par"])
            if pme_grid is not None:
                b1.pme_setup(pme_grid)
            with torch.cuda.stream(stream):
                b1.rebuild(pq, stream=stream)
            singles.append((b1, pq))
        f_tmp = torch.zeros_like(force)

        def timed_loop(fn, iters=20):
            evs = []
            with torch.cuda.stream(stream):
                for it in range(iters + 3):
                    if flush is not None:
                        flush.zero_()
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record(stream); fn(); b.record(stream)
                    if it >= 3:
                        evs.append((a, b))
            torch.cuda.synchronize()
            return sum(a.elapsed_time(b) for a, b in evs) / len(evs)

        t_two = timed_loop(lambda: [b1.step(pq, f_tmp, include_energy=True, graph=use_graph, stream=stream) for b1, pq in singles])
        t_fused = timed_loop(lambda: be.step(posq, force, posq_corr=corr, include_energy=True, graph=use_graph, stream=stream))
        # forces-only step (an integrator step that does not ask for the energy: the shared env-env pair ENERGIES are
        # skipped; u, W and the forces stay exact).  The reference always evaluates energies (ATMMetaForceImpl.cpp:113,116),
        # so the headline keeps them on; this is reported for information only.
        t_fonly = timed_loop(lambda: be.step(posq, force, posq_corr=corr, include_energy=False, graph=use_graph, stream=stream))
        two_sep = {"two_single_state_steps_ms": t_two, "fused_two_state_step_ms": t_fused, "speedup": t_two / t_fused,
                   "fused_forces_only_step_ms": t_fonly,
                   "note": "same kernels, same pair-list settings; the two single-state handles evaluate x and x+d separately"}
        for b1, _ in singles:
            b1.close()
        del singles, posq2, f_tmp

    # ---- Tier-1 kernels (copy-state, hybrid-force) against the HBM roofline, on a working set beyond L2
    tier1 = None
    if rank == 0 and world == 1 and not args.skip_tier1:
        with torch.cuda.stream(torch.cuda.default_stream(dev)):
            tier1 = tier1_hbm_probe(torch, atm, dev, flush)
        torch.cuda.empty_cache()

    # ---- end to end through the public call with HOST buffers: every step copies that step's coordinates H2D from
    #      pinned memory, runs the step, and reads forces + energies back D2H.  The replicas of the rank are split into
    #      --e2e-chunks handles on their own streams so that the copies of one chunk overlap the compute of another
    #      (everything still happens inside the timed step; PCIe is full duplex).
    e2e_ms = None
    h2d = d2h = 0
    e2e_chunks = 0
    sizes = []
    if R > 0 and not args.skip_e2e:
        sizes = e2e_split(args.e2e_split, R, args.e2e_chunks)
        e2e_chunks = len(sizes)
        bounds = [sum(sizes[:i]) for i in range(e2e_chunks + 1)]
        chunks = []
        for c in range(e2e_chunks):
            lo, hi = bounds[c], bounds[c + 1]
            if e2e_chunks == 1:
                bc = be
            else:
                bc = atm.ATMBackend(n, precision="mixed", num_replicas=hi - lo, device=local_rank)
                bc.set_displacements(s["displ"])
                bc.set_box(s["box"])
                for k in range(lo, hi):
                    bc.set_parameters(sched[replica_state[mine[k]]], replica=k - lo)
                bc.nb_setup(s["charge"], s["sigma"], s["epsilon"], s["cutoff"], s["ewald_alpha"], skin=args.skin,
                            skin_outer=args.skin_outer, exclusions=s["excl"], exception_pairs=s["exc14"],
                            exception_params=s["exc14_par"])
                if pme_grid is not None:
                    bc.pme_setup(pme_grid)
            chunks.append((bc, lo, hi))
        # the public host-buffer call (atm_host_pipeline_step, include/atm_b200.h): pinned coordinates in, pinned forces
        # and energy records out, the chunks' copies and kernels overlapped on streams of their own, one cached CUDA
        # graph launch per step; pair-list maintenance on the same cadence as the device-resident loop
        pipe = atm.HostPipeline([c[0] for c in chunks])
        en_h = torch.zeros((R, _capi.NUM_ENERGY_SLOTS), dtype=torch.float64).pin_memory()
        # the caller's pinned coordinate buffer: packed float3 (the charges are known to the back-end since nb_setup) or
        # OpenMM's float4 posq
        pos_host = (posq_h[:, :, :3].contiguous() if args.e2e_posq == "f3" else posq_h.clone()).pin_memory()
        pq_c = [pos_host[lo:hi] for _, lo, hi in chunks]
        f_c = [force_h[lo:hi] for _, lo, hi in chunks]
        en_c = [en_h[lo:hi] for _, lo, hi in chunks]
        # per-step jitter of the HOST coordinates: NJ pinned snapshots cycled through (the copy into the caller's pinned
        # buffer happens outside the event pair, like the integrator's position update it stands for)
        base_h = pos_host.clone()
        snaps = []
        rng_h = np.random.default_rng(3033 + rank)
        for j in range(4):
            z = base_h.clone()
            z[:, :n, :3] += torch.from_numpy(np.clip(rng_h.normal(0.0, JITTER_NM, (R, n, 3)), -2.5 * JITTER_NM, 2.5 * JITTER_NM).astype(np.float32))
            snaps.append(z)
        KE = min(K, 200)
        WE = 10   # the first prune step (k = 5) and its graph capture fall into the warm-up
        kinds_e = []
        ee = []
        with torch.cuda.stream(stream):
            pipe.step(pq_c, f_c, en_c, maintenance=pipe.REBUILD, stream=stream)   # first build: synchronous, verified
        schedule_e = [None] * (KE + WE)
        seen = {"plain": 0, "prune": 0, "rebuild": 0}
        for k in range(KE + WE):
            kd = "rebuild" if k % args.rebuild_every == 0 else ("prune" if k % args.prune_every == 0 else "plain")
            schedule_e[k] = kd
            if k >= WE:
                seen[kd] += 1
        for kd in ("prune", "rebuild", "plain"):      # kinds the window holds fewer than 3 times: sampled after it
            for _ in range(max(0, 3 - seen[kd])):
                schedule_e += [kd, "plain"]
        for k, kd in enumerate(schedule_e):
            pos_host.copy_(snaps[k % len(snaps)])
            with torch.cuda.stream(stream):
                if flush is not None:
                    flush.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(stream)
                # the prune runs BEFORE the step here whatever --prune-mode says: with several chunks on streams of their
                # own, side-stream prunes only add streams (measured: 0.155 vs 0.145 ms per step at 3 replicas)
                maint = {"rebuild": pipe.REBUILD, "prune": pipe.PRUNE, "plain": pipe.NONE}[kd]
                pipe.step(pq_c, f_c, en_c, maintenance=maint, stream=stream)
                b.record(stream)
                if k >= WE:
                    ee.append((kd, k < KE + WE, a, b))
            stream.synchronize()  # the step's result (forces, energies) is on the host before the next step starts
        pipe.close()
        torch.cuda.synchronize()
        e2e_comp = {}
        e2e_ms = 0.0
        f_e = {"rebuild": freq["rebuild"], "prune": freq["prune"], "plain": 1.0 - freq["rebuild"] - freq["prune"]}
        for kd in ("plain", "prune", "rebuild"):
            v = [a.elapsed_time(b) for kk, _, a, b in ee if kk == kd]
            t_kd = (float(np.median(v)) if len(v) < 50 else sum(v) / len(v)) if v else 0.0
            e2e_comp[kd] = {"ms": t_kd, "samples": len(v), "per_step": f_e[kd]}
            e2e_ms += t_kd * f_e[kd]
        e2e_window_ms = sum(a.elapsed_time(b) for _, inw, a, b in ee if inw) / KE
        h2d = pos_host.numel() * 4
        d2h = force_h.numel() * force_h.element_size() + R * _capi.NUM_ENERGY_SLOTS * 8
        for bc, *_ in chunks:
            if bc is not be:
                bc.close()
    if e2e_ms is None:
        e2e_comp, e2e_window_ms = {}, 0.0
    t = torch.tensor([e2e_ms or 0.0, e2e_window_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms, e2e_window_ms = float(t[0].item()), float(t[1].item())
    e2e_value = total_replicas * DT_FS * 1e-6 * 86400.0 / (e2e_ms * 1e-3) if e2e_ms > 0 else None

    if rank == 0:
        prop = torch.cuda.get_device_properties(dev)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        sm_max = (clocks or {}).get("sm_max_mhz") or peaks.get("sm_max_mhz") or 1965.0
        fp32_peak = prop.multi_processor_count * 128 * 2 * sm_max * 1e6 / 1e12
        pc, p1, p2 = stats_en[:R, 8].sum(), stats_en[:R, 9].sum(), stats_en[:R, 10].sum()
        pairs_two_state = 2.0 * pc + p1 + p2          # what two separate inner-context evaluations would compute
        pairs_computed = pc + p1 + p2
        roofline = None
        traffic = None
        traffic_src = None
        try:  # DRAM bytes of one nb2 launch: ncu cannot run inside a bench run, so this is the newest committed
            # `ncu --set full` capture of the same workload / replica count (profiles/r*_nb2_traffic.json), N = 1 only
            import glob
            for fn in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_nb2_traffic.json"))):
                tj = json.load(open(fn))
                if tj["workload"] == args.workload and tj["replicas"] == R and world == 1:
                    traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"]
                    traffic_src = os.path.relpath(fn, ROOT)
        except Exception:
            pass
        if nb2_ms:
            ach = pairs_two_state * FLOP_PER_PAIR / (nb2_ms * 1e-3) / 1e12
            roofline = {"kernel": "nb2_kernel (two-state direct space)", "bound": "fp32", "achieved": ach,
                        "peak": fp32_peak, "unit": "TFLOP/s", "frac": ach / fp32_peak, "traffic": traffic, "traffic_source": traffic_src,
                        "peak_source": "derived: SMs*128 lanes*2 flop*max SM clock (no measured fp32 peak in MEASURED_PEAKS.json)",
                        "flop_per_pair": FLOP_PER_PAIR, "pairs_two_state_equivalent": pairs_two_state,
                        "pairs_computed": pairs_computed, "nb2_ms": nb2_ms,
                        "computed_tflops": pairs_computed * FLOP_PER_PAIR / (nb2_ms * 1e-3) / 1e12,
                        "nb2_timed_over": "20 non-graph steps right after the timed region, CUDA events around the launch"}
            # the same launch against the HBM roofline (it is NOT the bound): algorithmic bytes = the pair list streamed
            # once (4 B per entry) + every site's coordinates and parameters read once (24 B) + the three force
            # accumulators written once (72 B per site)
            sites = float(nb_stats.get("sites", 0)) * R
            alg_bytes = 4.0 * float(nb_stats.get("list_entries", 0)) * R + sites * (24.0 + 72.0)
            hbm_peak = peaks.get("hbm_gbs", 6650.0)
            roofline["hbm"] = {"algorithmic_bytes": alg_bytes, "achieved": alg_bytes / (nb2_ms * 1e-3) / 1e9, "peak": hbm_peak,
                               "unit": "GB/s", "frac": alg_bytes / (nb2_ms * 1e-3) / 1e9 / hbm_peak,
                               "peak_source": "measured" if "hbm_gbs" in peaks else "fallback",
                               "note": "far below 1: the kernel is bound by FP32 instruction issue, not by memory"}
        cpu_baseline = None
        if world == 1:
            rate, sec, threads, cpu_done = cpu_oracle_rate(s, sched, args.cpu_steps)
            cpu_baseline = {"value": rate, "unit": "replica-ns/day", "cores": threads, "kind": "port",
                            "sample": f"{cpu_done} steps of 1 replica on the host cores ({sec:.3f} s/step, {cpu_done * sec:.1f} s of CPU work); replicas "
                                      f"run sequentially on the CPU so aggregate == per-replica rate"}
        line = {
            "metric": "aggregate replica-ns/day (ATM hot path)", "value": value, "unit": "replica-ns/day",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32 pair math, int64 fixed-point accumulation, f64 scalar stage",
            "data": "synthetic",
            "config": workload_config(args, label, s, max_per_rank, "device" if device_exchange else "host", use_graph, pme_grid,
                                      "none" if flush is None else ("flushed between steps, outside the per-step event pairs: 256 MiB memset" +
                                                                  (", then 256 MiB read (clean lines)" if args.flush_mode == "write+read" else ""))),
            "per_replica_ns_day": value / total_replicas, "us_per_replica_step": ms_per_step * 1e3 / max_per_rank,
            "ms_per_step_is": "steady state on the declared cadences: plain step + sum(component ms * per_step), every component "
                              "timed in this run with CUDA events (mean of the plain steps, median of the few maintenance / "
                              "exchange samples; max over ranks); window_ms_per_step is the raw K-step window",
            "components": comp, "window_ms_per_step": window_ms_per_step,
            "clocks": clocks, "gpu_launches": int(launches), "wall_s": wall,
            "e2e": {"value": e2e_value, "unit": "replica-ns/day", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": e2e_ms, "window_ms_per_step": e2e_window_ms, "components": e2e_comp, "chunks": e2e_chunks, "chunk_replicas": sizes,
                    "force_format": args.e2e_force, "posq_format": args.e2e_posq,
                    "call": "atm_host_pipeline_step (pinned host coordinates in, pinned host forces + energy records out, "
                            "one cached CUDA graph per step; pair-list maintenance on the bench cadence)"},
            "roofline": roofline, "cpu_baseline": cpu_baseline, "pair_list": nb_stats, "two_state_vs_two_separate": two_sep,
            "tier1_hbm_roofline": tier1,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


_RESULT_FD = None


def emit(line):
    """The ONE JSON line of the run, on the process's real stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)


def main():
    global _RESULT_FD
    args = parse_args()
    # stdout carries the result line and nothing else: whatever libraries print there (NCCL's version banner, ...) is sent
    # to stderr
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
