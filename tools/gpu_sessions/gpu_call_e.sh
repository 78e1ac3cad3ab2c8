#!/bin/bash
# one GPU-box session: host-pipeline tests, default bench, prune-kernel A/B (ATM_PRUNE_BLOCK=4 build), inner-skin / prune-cadence sweep
mkdir -p gpurun_out
PB4=$PWD/openmm-atmmetaforce-plugin_b200/libatm_b200_pb4.so
timeout 300 python -m pytest tests/test_gpu_host_pipeline.py -x -q > gpurun_out/e_pipe_tests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/e_pipe_tests.log
timeout 300 python bench.py > gpurun_out/e_bench_n1.json 2> gpurun_out/e_bench_n1.err; echo "bench rc=$?"
ATM_B200_LIB=$PB4 timeout 400 python -m pytest tests/test_gpu_nb2.py -x -q -k "abfe or rbfe or config3 or replicas" > gpurun_out/e_pb4_tests.log 2>&1; echo "pb4 pytest rc=$?"; tail -3 gpurun_out/e_pb4_tests.log
for R in 22 3; do
  for lib in base pb4; do
    if [ $lib = pb4 ]; then export ATM_B200_LIB=$PB4; else unset ATM_B200_LIB; fi
    echo "== time_step R=$R lib=$lib"; timeout 120 python tools/time_step.py --replicas $R --steps 30 --skin-outer 0.3 2>&1 | tail -1
  done
done
for R in 22 3; do
  for cfg in "base 0.1 10" "pb4 0.1 10" "pb4 0.07 7" "pb4 0.05 5" "base 0.05 5" "pb4 0.035 3"; do
    set -- $cfg
    if [ $1 = pb4 ]; then export ATM_B200_LIB=$PB4; else unset ATM_B200_LIB; fi
    timeout 120 python bench.py --replicas $R --skin $2 --prune-every $3 --e2e-chunks 3 --steps 400 --skip-two-separate --skip-tier1 --cpu-steps 1 > gpurun_out/e_sweep_R${R}_$1_skin$2.json 2>> gpurun_out/e_sweep.err; echo "R=$R $cfg rc=$?"
  done
done
unset ATM_B200_LIB
python - <<'P'
import json, glob
for f in sorted(glob.glob("gpurun_out/e_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["ms_per_step"], 4), round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), round(d["e2e"]["ms_per_step"], 4), "nb2", d["roofline"]["nb2_ms"])
    except Exception as e:
        print(f, "ERR", e)
P
