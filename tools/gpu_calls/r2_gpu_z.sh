#!/bin/bash
# round 2, call Z (1 GPU): last sanity of the final tree (smoke + a slice of the parity tests)
set -u
O=gpurun_out
mkdir -p $O
timeout 120 python __graft_entry__.py smoke > $O/r2z_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/r2z_smoke.log
timeout 150 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "packed_level_storage or device_resident_step or partitioned_on_one_gpu" > $O/r2z_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $O/r2z_pytest.log
