#!/bin/bash
# The GPU sessions of round 2, one stage per gpurun call:  gpurun -- bash tools/gpu_round2.sh <stage> [args]
# (small named recipes so that every number under profiles/ can be reproduced; tools/gpu_session.sh holds the generic
# building blocks -- ncu captures, launch lists, A/B of library builds -- that the stages call).
set -u
mkdir -p gpurun_out
stage=${1:-final}; shift || true
case "$stage" in
  final)   # whole GPU suite, smoke, the bench lines kept under profiles/, ncu captures of the PME kernels, sanitizer run
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
    ;;
  final_r2j)   # the same before the PME rewrite: suite, special-first A/B, nb2 / prune ncu captures, launch list, sanitizers
    timeout 1300 python -m pytest tests -m gpu -q > gpurun_out/gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/gpu_tests.log | cut -c1-250
    for sf in 0 1; do for r in 3 22; do ATM_B200_SPECIAL_FIRST=$sf python bench.py --steps 200 --warmup 50 --cpu-steps 1 --skip-two-separate --skip-tier1 --replicas $r > gpurun_out/r2j_sf${sf}_r$r.json 2>> gpurun_out/r2j.err; done; done
    bash tools/gpu_session.sh ncu_nb2 r2j
    bash tools/gpu_session.sh launches r2j
    bash tools/gpu_session.sh ncu_kernel r2j_prune nl_prune_kernel 2
    for tool in memcheck racecheck synccheck; do timeout 900 compute-sanitizer --tool $tool --kernel-regex kns=atm python tools/sanitize_case.py > gpurun_out/r2j_sanitizer_$tool.log 2>&1; echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY" gpurun_out/r2j_sanitizer_$tool.log | tail -2; done
    echo done
    ;;
  scaling)   # one 8-GPU box: the contract line at N = 1, 2, 4, 8 (both arms) and the multi-rank NCCL test; arg: tag
    tag=${1:-r2w}
    for n in 1 2 4 8; do
      if [ $n -eq 1 ]; then
        python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/${tag}_scale_n1.json 2> gpurun_out/${tag}_scale_n1.err
      else
        python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700 + n)) bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/${tag}_scale_n$n.json 2> gpurun_out/${tag}_scale_n$n.err
      fi
      echo "N=$n rc=$?"; cut -c1-260 gpurun_out/${tag}_scale_n$n.json | tail -1
    done
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29790 bench.py --impl reference --gpus 8 --steps 5 --warmup 1 > gpurun_out/${tag}_ref_n8.json 2> gpurun_out/${tag}_ref_n8.err; cut -c1-200 gpurun_out/${tag}_ref_n8.json
    timeout 600 python -m pytest tests/test_gpu_nccl.py -m gpu -q 2>&1 | tail -3
    ;;
  pme)   # PME parity tests, bench --pme, launch list
    timeout 900 python -m pytest tests/test_gpu_pme.py tests/test_gpu_facade.py tests/test_gpu_plugin.py -m gpu -q > gpurun_out/r2r_pme_tests.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/r2r_pme_tests.log | cut -c1-250
    for rep in 1 2; do for lib in libatm_b200.so; do
      ATM_B200_LIB=$PWD/openmm-atmmetaforce-plugin_b200/$lib timeout 600 python bench.py --pme --steps 100 --warmup 20 --cpu-steps 1 --skip-two-separate --skip-tier1 --skip-e2e > gpurun_out/r2r_pme_$lib.json 2> gpurun_out/r2r_pme_$lib.err
      python -c "
import json
j=json.loads(open('gpurun_out/r2r_pme_$lib.json').read().strip().splitlines()[-1])
print('$lib rep $rep ms_per_step', round(j['ms_per_step'],4), 'step', round(j['components']['step']['ms'],4))"
    done; done
    bash tools/gpu_session.sh launches r2r_pme --pme
    python - <<'PY'
import csv, collections
rows = list(csv.reader(l for l in open('gpurun_out/r2r_pme_bench_launches.csv') if l.startswith('"')))
hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value')
agg = collections.defaultdict(list)
for r in rows[2:]:
    try: agg[r[ki][:60]].append(float(r[vi].replace(',', '')))
    except Exception: pass
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    if 'at::' in k or 'nl_' in k or 'hrex' in k: continue
    v.sort(); print(f"{k:60s} n={len(v):4d} median={v[len(v)//2]/1000:9.1f} us")
PY
    ;;
  pme_ab)   # PME A/B: brick size of the spread, occupancy variants of the gather
    run() { # label env lib
      env $2 ATM_B200_LIB=$PWD/openmm-atmmetaforce-plugin_b200/$3 python bench.py --pme --steps 100 --warmup 20 --cpu-steps 1 --skip-two-separate --skip-tier1 --skip-e2e 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', 'ms_per_step', round(j['ms_per_step'],4), 'step', round(j['components']['step']['ms'],4))"
    }
    for rep in 1 2; do
      for t in 8 10 12 15 20; do run "tile=$t rep$rep" ATM_B200_PME_TILE=$t libatm_b200.so; done
      for lib in libatm_b200.so libatm_b200_g6.so libatm_b200_g8.so; do run "gather $lib rep$rep" X=1 $lib; done
    done
    ;;
  e2e)   # e2e leg: replicas per chunk of the host pipeline
    run() { # replicas split
      python bench.py --steps 100 --warmup 20 --cpu-steps 1 --skip-two-separate --skip-tier1 --replicas $1 --e2e-split $2 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=j['e2e']
print('R=$1 split=$2', 'device ms', round(j['ms_per_step'],4), 'e2e ms', round(e['ms_per_step'],4), 'plain', round(e['components']['plain']['ms'],4), 'value', round(e['value'],1))"
    }
    for sp in equal 2,4,5,5,4,2 2,4,8,6,2 1,3,7,7,3,1 2,6,8,6 1,2,4,8,5,2 1,2,16,2,1 2,9,9,2 11,11; do run 22 $sp; done
    for sp in equal 1,2,3,2,2,1 1,4,5,1 1,9,1 2,7,2; do run 11 $sp; done
    for sp in equal 1,4,1 2,2,2 3,3; do run 6 $sp; done
    for sp in equal 3 1,2; do run 3 $sp; done
    ;;
  cadence)   # suite + smoke of the final build, pair-list cadence sweep
    timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/gpu_tests.log | cut -c1-200
    timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-200
    run() { python bench.py --steps 200 --warmup 20 --cpu-steps 1 --skip-two-separate --skip-tier1 --skip-e2e --skin $1 --prune-every $2 --skin-outer $3 --rebuild-every $4 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); c=j['components']
print('skin $1 prune $2 outer $3 rebuild $4', 'ms_per_step', round(j['ms_per_step'],4), {k: round(v['ms'],4) for k,v in c.items()})"; }
    for p in "0.03 3" "0.04 4" "0.05 5" "0.06 6" "0.07 7" "0.08 8"; do run $p 0.3 40; done
    for o in "0.2 27" "0.25 33" "0.3 40" "0.35 47" "0.4 53"; do run 0.05 5 $o; done
    ;;
  *) echo "unknown stage $stage"; exit 2 ;;
esac
