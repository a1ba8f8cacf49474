#!/bin/bash
# round 2, call N (1 GPU): phase B edge body with the own node's mins hoisted out of the edge loop (6 fewer FSEL per
# edge and level) against the build before it, interleaved in one process; then the GPU test suite on the new build
set -u
O=gpurun_out
mkdir -p $O
timeout 400 python tools/ab_libs.py 1536x1204x70 build_ab/lib_before_hoist.so fesom2-accelerate_b200/lib/libfesom2-accelerate.so 5 > $O/r2n_ab_hoist_mid.log 2>&1; echo "ab mid rc=$?"; tail -3 $O/r2n_ab_hoist_mid.log
timeout 300 python tools/ab_libs.py 400x317x48 build_ab/lib_before_hoist.so fesom2-accelerate_b200/lib/libfesom2-accelerate.so 6 > $O/r2n_ab_hoist_core2.log 2>&1; echo "ab core2 rc=$?"; tail -3 $O/r2n_ab_hoist_core2.log
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r2n_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r2n_pytest.log
