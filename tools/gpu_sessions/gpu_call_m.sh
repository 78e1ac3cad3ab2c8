#!/bin/bash
mkdir -p gpurun_out
timeout 150 python -m pytest tests -x -q -m gpu > gpurun_out/m_gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/m_gpu_tests.log | cut -c1-300
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
