#!/bin/bash
# prune-kernel occupancy variants (cluster atoms in shared memory): ATM_PRUNE_BLOCK x ATM_PRUNE_MIN_BLOCKS
mkdir -p gpurun_out
D=$PWD/openmm-atmmetaforce-plugin_b200
for R in 22 3; do
  for v in p2m8 p4m6 p1m10 p2m12 p1m16; do
    export ATM_B200_LIB=$D/libatm_b200_$v.so
    echo "== time_step R=$R lib=$v $(timeout 120 python tools/time_step.py --replicas $R --steps 20 --skin-outer 0.3 2>&1 | tail -1 | cut -c1-110)"
  done
done 2>&1 | tee gpurun_out/g_prune_variants.log
