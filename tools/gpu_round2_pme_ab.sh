mkdir -p gpurun_out
run() { # label env lib
  env $2 ATM_B200_LIB=$PWD/openmm-atmmetaforce-plugin_b200/$3 python bench.py --pme --steps 100 --warmup 20 --cpu-steps 1 --skip-two-separate --skip-tier1 --skip-e2e 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', 'ms_per_step', round(j['ms_per_step'],4), 'step', round(j['components']['step']['ms'],4))"
}
for rep in 1 2; do
  for t in 8 10 12 15 20; do run "tile=$t rep$rep" ATM_B200_PME_TILE=$t libatm_b200.so; done
  for lib in libatm_b200.so libatm_b200_g6.so libatm_b200_g8.so; do run "gather $lib rep$rep" X=1 $lib; done
done
