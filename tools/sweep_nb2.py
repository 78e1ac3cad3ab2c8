"""Size sweep of the two-state direct-space kernel (BASELINE configs[4]): pure-water boxes of 10k-500k atoms with a
50-atom ligand, one replica, 0.9 nm cutoff.  Prints one JSON line per size: nb2 launch time (CUDA events around the
launch on the launching stream, L2 flushed between steps), pairs inside the cutoff, two-state-equivalent FP32 TFLOP/s
against the derived FP32 peak, useful-pair fraction, and the step / prune / rebuild times.
"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "openmm-atmmetaforce-plugin_b200", "python"))
import numpy as np
import torch
import atmmetaforce as atm
from atmmetaforce import synthetic

FLOP_PER_PAIR = 60.0
prop = torch.cuda.get_device_properties(0)
try:
    sm_max = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["sm_max_mhz"]
except Exception:
    sm_max = 1965.0
PEAK = prop.multi_processor_count * 128 * 2 * sm_max * 1e6 / 1e12
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
stream = torch.cuda.Stream()
sizes = [int(x) for x in sys.argv[1:]] or [10_000, 25_000, 50_000, 100_000, 250_000, 500_000]
for natoms in sizes:
    s = synthetic.water_box(natoms, n_lig=50)
    n = s["pos"].shape[0]
    be = atm.ATMBackend(n, precision="mixed", num_replicas=1)
    be.set_displacements(s["displ"])
    be.set_box(s["box"])
    be.set_parameters(synthetic.atm_schedule_22()[5])
    be.nb_setup(s["charge"], s["sigma"], s["epsilon"], s["cutoff"], s["ewald_alpha"], skin=0.1, skin_outer=0.3, exclusions=s["excl"])
    posq = np.zeros((1, be.P, 4), np.float32)
    posq[0, :n, :3] = s["pos"]; posq[0, :n, 3] = s["charge"]
    posq = torch.from_numpy(posq).cuda()
    force = torch.zeros((1, 3 * be.P), dtype=torch.int64, device="cuda")

    def timed(fn, iters):
        evs = []
        with torch.cuda.stream(stream):
            for it in range(iters + 3):
                flush.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(stream); fn(); b.record(stream)
                if it >= 3:
                    evs.append((a, b))
        torch.cuda.synchronize()
        return float(np.median([a.elapsed_time(b) for a, b in evs]))

    with torch.cuda.stream(stream):
        be.rebuild(posq, stream=stream)
        be.step(posq, force, collect_stats=True, stream=stream)
    en = be.get_energies(stream=stream)[0]
    pc, p1, p2 = en[8], en[9], en[10]
    ms_rebuild = timed(lambda: be.rebuild(posq, stream=stream), 5)
    ms_prune = timed(lambda: be.prune(posq, stream=stream), 10)
    ms_step = timed(lambda: be.step(posq, force, graph=True, stream=stream), 30)
    be.profile_enable(True)
    with torch.cuda.stream(stream):
        for _ in range(20):
            flush.zero_()
            be.step(posq, force, stream=stream)
    tot, cnt = be.profile_read()
    be.profile_enable(False)
    nb2_ms = tot / cnt
    st = be.nb_stats()
    two_state = 2.0 * pc + p1 + p2
    tf = two_state * FLOP_PER_PAIR / (nb2_ms * 1e-3) / 1e12
    print(json.dumps({"atoms": n, "nb2_ms": nb2_ms, "step_ms": ms_step, "prune_ms": ms_prune, "rebuild_ms": ms_rebuild,
                      "pairs_in_cutoff": pc + p1 + p2, "pairs_two_state_equivalent": two_state,
                      "useful_pair_fraction": (pc + p1 + p2) / (st["list_entries"] * 8.0),
                      "tflops_two_state_equivalent": tf, "frac_of_fp32_peak": tf / PEAK, "fp32_peak_tflops": PEAK,
                      "tflops_computed": (pc + p1 + p2) * FLOP_PER_PAIR / (nb2_ms * 1e-3) / 1e12,
                      "ns_per_day_at_1fs": 86.4 / ms_step, "l2": "flushed"}), flush=True)
    be.close()
    del posq, force
    torch.cuda.empty_cache()
