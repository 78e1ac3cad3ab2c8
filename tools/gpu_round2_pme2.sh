#!/bin/bash
# PME session 2: parity tests, A/B of the gather occupancy variants, launch list, ncu full of spread / gather.
mkdir -p gpurun_out
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
