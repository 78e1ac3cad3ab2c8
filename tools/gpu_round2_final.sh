mkdir -p gpurun_out
timeout 1300 python -m pytest tests -m gpu -q > gpurun_out/gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/gpu_tests.log | cut -c1-250
for sf in 0 1; do for r in 3 22; do ATM_B200_SPECIAL_FIRST=$sf python bench.py --steps 200 --warmup 50 --cpu-steps 1 --skip-two-separate --skip-tier1 --replicas $r > gpurun_out/r2j_sf${sf}_r$r.json 2>> gpurun_out/r2j.err; done; done
bash tools/gpu_session.sh ncu_nb2 r2j
bash tools/gpu_session.sh launches r2j
bash tools/gpu_session.sh ncu_kernel r2j_prune nl_prune_kernel 2
for tool in memcheck racecheck synccheck; do timeout 900 compute-sanitizer --tool $tool --kernel-regex kns=atm python tools/sanitize_case.py > gpurun_out/r2j_sanitizer_$tool.log 2>&1; echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY" gpurun_out/r2j_sanitizer_$tool.log | tail -2; done
echo done
