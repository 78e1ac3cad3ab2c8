#!/bin/bash
# usage: gpu_call_scale.sh N   -- the contract's multi-GPU launch on N GPUs of one box
N=$1
mkdir -p gpurun_out
timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 1000 --warmup 20 > gpurun_out/s_bench_n$N.json 2> gpurun_out/s_bench_n$N.err; echo "bench N=$N rc=$?"
tail -2 gpurun_out/s_bench_n$N.err
python - <<P
import json
d = json.loads(open("gpurun_out/s_bench_n$N.json").read().strip().splitlines()[-1])
print("N=$N", round(d["ms_per_step"], 4), round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), round(d["e2e"]["ms_per_step"], 4), d["clocks"])
P
