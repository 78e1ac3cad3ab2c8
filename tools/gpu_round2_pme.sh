#!/bin/bash
# PME session: parity tests of the float mesh pipeline, A/B against the double pipeline, per-kernel launch list.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pme.py tests/test_gpu_facade.py tests/test_gpu_plugin.py -m gpu -q > gpurun_out/r2q_pme_tests.log 2>&1; echo "pytest rc=$?"; tail -40 gpurun_out/r2q_pme_tests.log | cut -c1-250
for f64 in 0 1; do
  ATM_B200_PME_F64=$f64 timeout 600 python bench.py --pme --steps 100 --warmup 20 --cpu-steps 1 --skip-two-separate --skip-tier1 --skip-e2e > gpurun_out/r2q_pme_f64_$f64.json 2> gpurun_out/r2q_pme_f64_$f64.err
  echo "f64=$f64 rc=$?"; python -c "
import json,sys
j=json.loads(open('gpurun_out/r2q_pme_f64_$f64.json').read().strip().splitlines()[-1])
print('ms_per_step', j['ms_per_step'], {k:v['ms'] for k,v in j['components'].items()})"
done
bash tools/gpu_session.sh launches r2q_pme --pme
python - <<'PY'
import csv, collections
rows = list(csv.reader(l for l in open('gpurun_out/r2q_pme_bench_launches.csv') if l.startswith('"')))
hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value')
agg = collections.defaultdict(list)
for r in rows[2:]:
    try: agg[r[ki][:60]].append(float(r[vi].replace(',', '')))
    except Exception: pass
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    v.sort(); print(f"{k:60s} n={len(v):4d} median={v[len(v)//2]/1000:9.1f} us")
PY
bash tools/gpu_session.sh ncu_kernel r2q_pme_spread pme_spread_tile 6 --pme
bash tools/gpu_session.sh ncu_kernel r2q_pme_gather pme_gather_f 6 --pme
