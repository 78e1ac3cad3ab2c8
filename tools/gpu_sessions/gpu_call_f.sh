#!/bin/bash
# one GPU-box session: full GPU test suite on the new prune kernel, prune timing A/B (4 vs 8 lists per block), skin / cadence sweep
mkdir -p gpurun_out
PW8=$PWD/openmm-atmmetaforce-plugin_b200/libatm_b200_pw8.so
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/f_gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/f_gpu_tests.log
for R in 22 3; do
  for lib in main pw8; do
    if [ $lib = pw8 ]; then export ATM_B200_LIB=$PW8; else unset ATM_B200_LIB; fi
    echo "== time_step R=$R lib=$lib"; timeout 120 python tools/time_step.py --replicas $R --steps 30 --skin-outer 0.3 2>&1 | tail -1
  done
done
for R in 22 3; do
  for cfg in "main 0.1 10" "main 0.07 7" "main 0.05 5" "pw8 0.05 5" "main 0.04 4"; do
    set -- $cfg
    if [ $1 = pw8 ]; then export ATM_B200_LIB=$PW8; else unset ATM_B200_LIB; fi
    timeout 120 python bench.py --replicas $R --skin $2 --prune-every $3 --e2e-chunks 3 --steps 400 --skip-two-separate --skip-tier1 --cpu-steps 1 > gpurun_out/f_sweep_R${R}_$1_skin$2.json 2>> gpurun_out/f_sweep.err; echo "R=$R $cfg rc=$?"
  done
done
unset ATM_B200_LIB
python - <<'P'
import json, glob
for f in sorted(glob.glob("gpurun_out/f_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["ms_per_step"], 4), round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), round(d["e2e"]["ms_per_step"], 4), "nb2", d["roofline"]["nb2_ms"])
    except Exception as e:
        print(f, "ERR", e)
P
