#!/bin/bash
# Final session of round 2 (after the PME rewrite and the external-force hook): whole GPU suite, smoke, the bench lines
# kept under profiles/, ncu captures of the PME kernels, sanitizer run.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/gpu_tests.log | cut -c1-250
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2t_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2t_smoke.log | cut -c1-200
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2t_bench_driver.json 2> gpurun_out/r2t_bench_driver.err; echo "driver bench rc=$?"
timeout 900 python bench.py > gpurun_out/r2t_bench_default.json 2> gpurun_out/r2t_bench_default.err; echo "default bench rc=$?"
timeout 600 python bench.py --pme --steps 100 --warmup 20 > gpurun_out/r2t_bench_pme.json 2> gpurun_out/r2t_bench_pme.err; echo "pme bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2t_bench_ref.json 2> gpurun_out/r2t_bench_ref.err; echo "reference arm rc=$?"
python - <<'PY'
import json
for t in ("driver", "default", "pme", "ref"):
    try:
        j = json.loads(open(f"gpurun_out/r2t_bench_{t}.json").read().strip().splitlines()[-1])
        print(t, "ms/step", round(j["ms_per_step"], 4), "value", round(j["value"], 1), "e2e", (j.get("e2e") or {}).get("value"), "frac", (j.get("roofline") or {}).get("frac"))
    except Exception as e:
        print(t, "unreadable:", e)
PY
bash tools/gpu_session.sh launches r2t_pme --pme
bash tools/gpu_session.sh ncu_kernel r2t_pme_gather pme_gather_f 6 --pme
bash tools/gpu_session.sh ncu_kernel r2t_pme_convolve pme_convolve_f 6 --pme
bash tools/gpu_session.sh ncu_kernel r2t_pme_spread pme_spread_tile 6 --pme
for tool in memcheck racecheck synccheck; do timeout 900 compute-sanitizer --tool $tool --kernel-regex kns=atm python tools/sanitize_case.py > gpurun_out/r2t_sanitizer_$tool.log 2>&1; echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY" gpurun_out/r2t_sanitizer_$tool.log | tail -2; done
echo done
