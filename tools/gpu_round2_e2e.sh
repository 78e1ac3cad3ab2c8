#!/bin/bash
# e2e leg: how the replicas of a rank are split over the chunks of the host pipeline
mkdir -p gpurun_out
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
