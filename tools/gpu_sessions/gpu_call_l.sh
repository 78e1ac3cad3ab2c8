#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_nb2.py tests/test_gpu_host_pipeline.py -x -q > gpurun_out/l_gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -1 gpurun_out/l_gpu_tests.log
timeout 400 python bench.py > gpurun_out/l_bench_n1.json 2> gpurun_out/l_bench_n1.err; echo "bench rc=$?"
timeout 120 python bench.py --replicas 3 --e2e-chunks 3 --steps 600 --skip-two-separate --skip-tier1 --cpu-steps 1 > gpurun_out/l_bench_R3.json 2>> gpurun_out/l.err
timeout 120 python bench.py --replicas 6 --e2e-chunks 3 --steps 600 --skip-two-separate --skip-tier1 --cpu-steps 1 > gpurun_out/l_bench_R6.json 2>> gpurun_out/l.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:nb2_kernel -s 5 -c 1 -f -o gpurun_out/l_nb2_full python bench.py --steps 4 --warmup 3 --cpu-steps 1 --skip-two-separate --skip-tier1 > gpurun_out/l_ncu_nb2.log 2>&1; echo "ncu nb2 rc=$?"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/l_bench_launches.csv python bench.py --steps 40 --warmup 3 --cpu-steps 1 --skip-tier1 > gpurun_out/l_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
python - <<'P'
import json, glob
for f in sorted(glob.glob("gpurun_out/l_bench_*.json")):
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, round(d["ms_per_step"], 4), round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), round(d["e2e"]["ms_per_step"], 4), "nb2", d["roofline"]["nb2_ms"], "frac", round(d["roofline"]["frac"], 3))
P
