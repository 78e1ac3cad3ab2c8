"""bench.py contract checks that need no GPU: the reference arm (the oracle port on the host cores) prints ONE JSON line
with the keys the driver reads, on the same metric / unit / config as the GPU arm."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    env = dict(os.environ, OMP_NUM_THREADS="1")   # what torchrun exports to its workers
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=900, cwd=ROOT, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "replica-ns/day" and d["higher_is_better"] is True
    assert d["metric"] == "aggregate replica-ns/day (ATM hot path)"
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 1 and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["vs_baseline"] is None and d["data"] == "synthetic"
    assert "BASELINE configs[2]" in d["config"]["workload"] and d["config"]["replicas"] == 22
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # same config keys as the GPU arm prints (both go through bench.workload_config)
    sys.path.insert(0, ROOT)
    import bench
    import inspect
    keys = set(d["config"])
    assert {"workload", "replicas", "replicas_per_rank", "atoms", "dt_fs", "cutoff_nm", "skin_nm", "skin_outer_nm", "prune_every",
            "rebuild_every", "exchange_every", "prune_mode", "exchange", "cuda_graph", "pme_grid", "l2", "jitter_nm"} == keys
    assert "workload_config(args, label, s, max_per_rank" in inspect.getsource(bench.run_b200)
    # what was timed is what is printed: a step = every replica once (or a stated sample of them), threads = the cores
    # this process may use even when the launcher exported OMP_NUM_THREADS=1
    assert d["replicas_per_step_timed"] == 22 and abs(d["ms_per_step_all_replicas"] - d["ms_per_step"]) < 1e-9
    assert cb["cores"] == len(os.sched_getaffinity(0))
    assert abs(d["value"] - 22 * 86400.0 * 1e-6 / (d["ms_per_step"] * 1e-3)) < 1e-9 * d["value"]


def test_bench_defaults_are_the_documented_workload():
    sys.path.insert(0, ROOT)
    import importlib
    bench = importlib.import_module("bench")
    argv, sys.argv = sys.argv, ["bench.py"]
    try:
        a = bench.parse_args()
    finally:
        sys.argv = argv
    assert (a.gpus, a.impl, a.workload, a.replicas) == (1, "b200", "config3", 22)
    assert a.warmup >= 3 and a.steps >= 100
    assert (a.skin, a.skin_outer, a.prune_every, a.rebuild_every) == (0.05, 0.3, 5, 40)   # DESIGN.md "Inner skin and prune cadence"
