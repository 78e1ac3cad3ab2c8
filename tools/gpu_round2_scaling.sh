#!/bin/bash
# One 8-GPU box: the contract's launch line at N = 1, 2, 4, 8 back to back (both arms), outputs under gpurun_out/.
mkdir -p gpurun_out
tag=${1:-r2k}
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
