#!/bin/bash
# One gpurun call: `gpurun -- bash tools/gpu_session.sh <stage> [args]`.  Every stage writes under gpurun_out/.
# Stages are small and named so that the sequence of GPU sessions of a round is reproducible from the history.
set -u
mkdir -p gpurun_out
stage=${1:-tests}; shift || true
case "$stage" in
  tests)      # the whole GPU suite
    timeout 1500 python -m pytest tests -m gpu -x -q "$@" > gpurun_out/gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/gpu_tests.log | cut -c1-300 ;;
  ab)         # A/B of several builds of the library on one box: ab <tag> <libA,libB,...> [bench args]
    tag=$1; libs=${2//,/ }; shift 2
    rm -f gpurun_out/ab_$tag.log
    for rep in 1 2; do for lib in $libs; do
      echo "== $lib (rep $rep)" >> gpurun_out/ab_$tag.log
      ATM_B200_LIB=$PWD/openmm-atmmetaforce-plugin_b200/$lib timeout 600 python bench.py --steps 200 --warmup 10 --cpu-steps 1 --skip-two-separate --skip-tier1 "$@" >> gpurun_out/ab_$tag.log 2>&1
    done; done
    python - "$tag" <<'PY'
import json, sys
tag = sys.argv[1]
cur = None
for ln in open(f"gpurun_out/ab_{tag}.log"):
    if ln.startswith("== "):
        cur = ln.strip()
    elif ln.startswith("{"):
        j = json.loads(ln)
        r = j.get("roofline") or {}
        print(cur, "ms/step", round(j["ms_per_step"], 4), "nb2_ms", round(r.get("nb2_ms") or 0, 4), "frac", round(r.get("frac") or 0, 3), "e2e_ms", round(j["e2e"]["ms_per_step"], 4))
PY
    ;;
  ncu_nb2)    # ncu --set full of one steady-state nb2 launch: ncu_nb2 <tag> [bench args]
    tag=$1; shift
    timeout 900 ncu --set full --clock-control none --import-source on -k regex:nb2_kernel -s 5 -c 1 -f -o gpurun_out/${tag}_nb2_full \
      python bench.py --steps 4 --warmup 3 --cpu-steps 1 --skip-two-separate --skip-tier1 "$@" > gpurun_out/${tag}_nb2_full.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/${tag}_nb2_full.log | cut -c1-200 ;;
  ncu_kernel) # ncu --set full of one launch of any kernel: ncu_kernel <tag> <regex> <skip> [bench args]
    tag=$1; rx=$2; skip=$3; shift 3
    timeout 900 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -f -o gpurun_out/${tag}_full \
      python bench.py --steps 4 --warmup 3 --cpu-steps 1 --skip-two-separate --skip-tier1 "$@" > gpurun_out/${tag}_full.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/${tag}_full.log | cut -c1-200 ;;
  launches)   # ncu launch list of a bench run: launches <tag> [bench args]
    tag=$1; shift
    timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/${tag}_bench_launches.csv \
      python bench.py --steps 40 --warmup 3 --cpu-steps 1 --skip-tier1 --skip-two-separate "$@" > gpurun_out/${tag}_bench_launches.log 2>&1; echo "ncu rc=$?" ;;
  bench)      # plain bench line(s): bench <tag> [bench args]
    tag=$1; shift
    timeout 900 python bench.py "$@" > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"; cut -c1-600 gpurun_out/${tag}_bench.json ;;
  *) echo "unknown stage $stage"; exit 2 ;;
esac
