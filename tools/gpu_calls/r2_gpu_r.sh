#!/bin/bash
# round 2, call R (1 GPU): phase A with area_inv pulled into L2 one item ahead (WT_FLAGS bit 2)
set -u
O=gpurun_out
mkdir -p $O
CFG="WT_FLAGS=0;WT_FLAGS=2"
timeout 150 python tools/ab_knobs.py 60x52x48 "$CFG" 1 > $O/r2r_ab_small.log 2>&1; rc=$?; echo "ab small rc=$rc"; tail -2 $O/r2r_ab_small.log
[ $rc -eq 0 ] || exit 1
timeout 400 python tools/ab_knobs.py 1536x1204x70 "$CFG" 6 > $O/r2r_ab_mid.log 2>&1; echo "ab mid rc=$?"; tail -2 $O/r2r_ab_mid.log
timeout 300 python tools/ab_knobs.py 400x317x48 "$CFG" 8 > $O/r2r_ab_core2.log 2>&1; echo "ab core2 rc=$?"; tail -2 $O/r2r_ab_core2.log
