#!/bin/bash
# round 2, call Q (1 GPU): phase B with the branch-free b3 vertical + min_one against the build before it; GPU tests
set -u
O=gpurun_out
mkdir -p $O
timeout 400 python tools/ab_libs.py 1536x1204x70 build_ab/lib_before_dtarea.so fesom2-accelerate_b200/lib/libfesom2-accelerate.so 5 > $O/r2q_ab_b3v_mid.log 2>&1; echo "ab mid rc=$?"; tail -3 $O/r2q_ab_b3v_mid.log
timeout 300 python tools/ab_libs.py 400x317x48 build_ab/lib_before_dtarea.so fesom2-accelerate_b200/lib/libfesom2-accelerate.so 6 > $O/r2q_ab_b3v_core2.log 2>&1; echo "ab core2 rc=$?"; tail -3 $O/r2q_ab_b3v_core2.log
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r2q_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r2q_pytest.log
