#!/bin/bash
# bench lines of the other BASELINE configs on one B200 (r1d code): rbfe fixture (1 replica), abfe fixture (1 replica), config4 (22 replicas x 100k atoms)
mkdir -p gpurun_out
timeout 200 python bench.py --workload rbfe --replicas 1 --e2e-chunks 1 --skip-tier1 --cpu-steps 20 > gpurun_out/i_bench_rbfe.json 2> gpurun_out/i_bench_rbfe.err; echo "rbfe rc=$?"
timeout 200 python bench.py --workload abfe --replicas 1 --e2e-chunks 1 --skip-tier1 --cpu-steps 20 > gpurun_out/i_bench_abfe.json 2> gpurun_out/i_bench_abfe.err; echo "abfe rc=$?"
timeout 400 python bench.py --workload config4 --steps 300 --skip-tier1 --cpu-steps 5 > gpurun_out/i_bench_config4.json 2> gpurun_out/i_bench_config4.err; echo "config4 rc=$?"
timeout 200 python bench.py --dt-fs 4 --steps 400 --skip-tier1 --skip-two-separate --cpu-steps 1 > gpurun_out/i_bench_dt4.json 2> gpurun_out/i_bench_dt4.err; echo "dt4 rc=$?"
python - <<'P'
import json, glob
for f in sorted(glob.glob("gpurun_out/i_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["ms_per_step"], 4), round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), round(d["e2e"]["ms_per_step"], 4), "nb2", d["roofline"]["nb2_ms"], "frac", round(d["roofline"]["frac"], 3), "2sep", (d.get("two_state_vs_two_separate") or {}).get("speedup"), "cpu", d["cpu_baseline"]["value"])
    except Exception as e:
        print(f, "ERR", e)
P
tail -3 gpurun_out/i_bench_*.err
