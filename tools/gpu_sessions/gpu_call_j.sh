#!/bin/bash
# A/B of the nb2 inner-loop change (one select on r^2, five-instruction Coulomb): full GPU suite on the experimental build, then bench main vs experimental
mkdir -p gpurun_out
SEL=$PWD/openmm-atmmetaforce-plugin_b200/libatm_b200_sel.so
ATM_B200_LIB=$SEL timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/j_gpu_tests_sel.log 2>&1; echo "pytest(sel) rc=$?"; tail -2 gpurun_out/j_gpu_tests_sel.log
for R in 22 3; do
  for lib in main sel main sel; do
    if [ $lib = sel ]; then export ATM_B200_LIB=$SEL; else unset ATM_B200_LIB; fi
    timeout 120 python bench.py --replicas $R --e2e-chunks 3 --steps 600 --skip-two-separate --skip-tier1 --cpu-steps 1 > gpurun_out/j_tmp.json 2>> gpurun_out/j_sweep.err
    python - <<P
import json
d=json.loads(open("gpurun_out/j_tmp.json").read().strip().splitlines()[-1])
print("R=$R lib=$lib ms/step %.4f nb2 %.4f frac %.3f e2e_ms %.4f" % (d["ms_per_step"], d["roofline"]["nb2_ms"], d["roofline"]["frac"], d["e2e"]["ms_per_step"]))
P
  done
done 2>&1 | tee gpurun_out/j_ab.log
