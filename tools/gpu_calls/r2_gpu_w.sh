#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -rA > $O/r2w_pytest_multi.log 2>&1; echo "pytest multi rc=$?"; tail -6 $O/r2w_pytest_multi.log
