#!/bin/bash
# round 2, call P (1 GPU): cached dt/area in phase B (knob DT_AREA) and the L2 prefetch of the c-vertical operands (WT_FLAGS 1 = off)
set -u
O=gpurun_out
mkdir -p $O
CFG="DT_AREA=0;DT_AREA=1;DT_AREA=1,WT_FLAGS=1;DT_AREA=0,WT_FLAGS=1"
timeout 150 python tools/ab_knobs.py 60x52x48 "$CFG" 1 > $O/r2p_ab_small.log 2>&1; rc=$?; echo "ab small rc=$rc"; tail -4 $O/r2p_ab_small.log
[ $rc -eq 0 ] || exit 1
timeout 500 python tools/ab_knobs.py 1536x1204x70 "$CFG" 5 > $O/r2p_ab_mid.log 2>&1; echo "ab mid rc=$?"; tail -4 $O/r2p_ab_mid.log
timeout 300 python tools/ab_knobs.py 400x317x48 "$CFG" 6 > $O/r2p_ab_core2.log 2>&1; echo "ab core2 rc=$?"; tail -4 $O/r2p_ab_core2.log
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r2p_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r2p_pytest.log
