#!/bin/bash
# one GPU-box session: host-pipeline tests, the default bench line, e2e chunk sweep at the per-rank replica counts of N = 1/4/8
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_host_pipeline.py -x -q > gpurun_out/d_pipe_tests.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/d_pipe_tests.log
timeout 300 python bench.py > gpurun_out/d_bench_n1.json 2> gpurun_out/d_bench_n1.err; echo "bench rc=$?"
for cfg in "22 11" "6 6" "6 3" "3 3" "3 1"; do
  set -- $cfg
  timeout 120 python bench.py --replicas $1 --e2e-chunks $2 --steps 200 --skip-two-separate --skip-tier1 --cpu-steps 1 > gpurun_out/d_bench_r$1_c$2.json 2>> gpurun_out/d_bench_sweep.err; echo "R=$1 chunks=$2 rc=$?"
done
python - <<'P'
import json, glob
for f in sorted(glob.glob("gpurun_out/d_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["ms_per_step"], 4), round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), round(d["e2e"]["ms_per_step"], 4))
    except Exception as e:
        print(f, "ERR", e)
P
