#!/bin/bash
# round 2, call X (1 GPU): when the tile-level prefetch is issued on the two-stage ring (two tile periods ahead vs one)
set -u
O=gpurun_out
mkdir -p $O
timeout 150 python tools/ab_knobs.py 96x74x70 "WT_OPT=6;WT_OPT=134;WT_OPT=390" 1 > $O/r2x_ab_small70.log 2>&1; rc=$?; echo "ab small70 rc=$rc"; tail -3 $O/r2x_ab_small70.log
[ $rc -eq 0 ] || exit 1
timeout 400 python tools/ab_knobs.py 1536x1204x70 "WT_OPT=6;WT_OPT=134;WT_OPT=390" 6 > $O/r2x_ab_mid.log 2>&1; echo "ab mid rc=$?"; tail -3 $O/r2x_ab_mid.log
