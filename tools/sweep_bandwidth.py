"""HBM roofline of the two Tier-1 bandwidth kernels (CopyState, HybridForce) over atom counts (BASELINE configs[4]
extended to 1-16 M atoms so that the working set exceeds the 126 MB L2).  Prints one JSON line per (kernel, N).

Algorithmic bytes: copy 112*N (mixed) / 64*N (single); merge 96*N.  Peak = MEASURED_PEAKS.json hbm_gbs (copy bandwidth).
"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "openmm-atmmetaforce-plugin_b200", "python"))
import numpy as np
import torch
import atmmetaforce as atm

try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    SRC = "measured"
except Exception:
    PEAK, SRC = 6650.0, "fallback"

flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")


def timed(fn, iters=20):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts)), float(np.min(ts))


SIZES = [int(x) for x in sys.argv[1:]] or [25_000, 100_000, 500_000, 1_000_000, 4_000_000, 8_000_000, 16_000_000]
for n in SIZES:
    P = 32 * ((n + 31) // 32)
    d = np.zeros((n, 3)); d[:50] = [2.2, 2.2, 2.2]
    for mode, bpa in (("mixed", 112), ("single", 64)):
        be = atm.ATMBackend(n, padded_num_particles=P, precision=mode)
        be.set_displacements(d)
        posq = torch.rand((P, 4), device="cuda")
        p1, p2 = torch.empty_like(posq), torch.empty_like(posq)
        if mode == "mixed":
            c, c1, c2 = torch.zeros_like(posq), torch.empty_like(posq), torch.empty_like(posq)
            fn = lambda: be.copy_state(posq, p1, p2, c, c1, c2)
        else:
            fn = lambda: be.copy_state(posq, p1, p2)
        med, best = timed(fn)
        gbs = bpa * n / (med * 1e-3) / 1e9
        print(json.dumps({"kernel": f"copy_state[{mode}]", "atoms": n, "bytes": bpa * n, "ms_median": med, "ms_min": best,
                          "GBps": gbs, "frac_of_peak": gbs / PEAK, "peak": PEAK, "peak_source": SRC, "l2": "flushed"}))
        if mode == "mixed":
            f0 = torch.zeros(3 * P, dtype=torch.int64, device="cuda")
            f1 = torch.randint(-2**40, 2**40, (3 * P,), dtype=torch.int64, device="cuda")
            f2 = torch.randint(-2**40, 2**40, (3 * P,), dtype=torch.int64, device="cuda")
            med, best = timed(lambda: be.hybrid_force(f0, f1, f2, 0.37))
            gbs = 96 * n / (med * 1e-3) / 1e9
            print(json.dumps({"kernel": "hybrid_force", "atoms": n, "bytes": 96 * n, "ms_median": med, "ms_min": best,
                              "GBps": gbs, "frac_of_peak": gbs / PEAK, "peak": PEAK, "peak_source": SRC, "l2": "flushed"}))
            del f0, f1, f2
        be.close()
        del posq, p1, p2
    torch.cuda.empty_cache()
