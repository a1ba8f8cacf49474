#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r2k_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r2k_pytest.log
bash tools/gpu_calls/r2_gpu_j.sh
