#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r2t_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r2t_pytest.log
