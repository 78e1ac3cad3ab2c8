#!/bin/bash
# validation + measurement session of the r1d state: full GPU suite, smoke, default bench, ncu launch list, ncu full sets, sanitizer
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/h_gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/h_gpu_tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/h_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/h_smoke.log
timeout 400 python bench.py > gpurun_out/h_bench_n1.json 2> gpurun_out/h_bench_n1.err; echo "bench rc=$?"
for cfg in "6 3" "3 3"; do set -- $cfg
  timeout 120 python bench.py --replicas $1 --e2e-chunks $2 --steps 400 --skip-two-separate --skip-tier1 --cpu-steps 1 > gpurun_out/h_bench_R$1.json 2>> gpurun_out/h_sweep.err; echo "R=$1 rc=$?"
done
timeout 120 python tools/time_step.py --replicas 22 --steps 30 --skin 0.05 --skin-outer 0.3 2>&1 | tail -1 > gpurun_out/h_time_step.log
timeout 120 python tools/time_step.py --replicas 3 --steps 30 --skin 0.05 --skin-outer 0.3 2>&1 | tail -1 >> gpurun_out/h_time_step.log
cat gpurun_out/h_time_step.log
# ncu: launch list of a short bench run (shares), then one full-set capture each of nb2 and the prune kernel
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/h_bench_launches.csv python bench.py --steps 40 --warmup 3 --cpu-steps 1 --skip-tier1 > gpurun_out/h_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:nb2_kernel -s 5 -c 1 -f -o gpurun_out/h_nb2_full python bench.py --steps 4 --warmup 3 --cpu-steps 1 --skip-two-separate --skip-tier1 > gpurun_out/h_ncu_nb2.log 2>&1; echo "ncu nb2 rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:nl_prune_kernel -s 2 -c 1 -f -o gpurun_out/h_prune_full python bench.py --steps 12 --warmup 3 --cpu-steps 1 --skip-two-separate --skip-tier1 > gpurun_out/h_ncu_prune.log 2>&1; echo "ncu prune rc=$?"
for tool in memcheck racecheck synccheck; do
  timeout 500 compute-sanitizer --tool $tool --kernel-regex kns=atm python tools/sanitize_case.py > gpurun_out/h_sanitizer_$tool.log 2>&1; echo "$tool rc=$?"; grep -E "SUMMARY|pipeline u" gpurun_out/h_sanitizer_$tool.log | tail -2
done
python - <<'P'
import json, glob
for f in sorted(glob.glob("gpurun_out/h_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["ms_per_step"], 4), round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), round(d["e2e"]["ms_per_step"], 4), "nb2", d["roofline"]["nb2_ms"], "frac", round(d["roofline"]["frac"], 3))
    except Exception as e:
        print(f, "ERR", e)
P
ls -la gpurun_out/*.ncu-rep
