#!/bin/bash
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_facade.py -x -q > gpurun_out/n_gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/n_gpu_tests.log | cut -c1-400
