#!/bin/bash
# round 2, call L (1 GPU): the a1 pass folded into the consumers' edge loop (no converter warps, no pass over the rows)
set -u
O=gpurun_out
mkdir -p $O
CFG="WT_CONV=2;WT_CONV=-1;WT_CONV=-1,WT_ISSUERS=2;WT_CONV=-1,WT_OPT=14"
timeout 150 python tools/ab_knobs.py 96x74x70 "$CFG" 1 > $O/r2l_ab_small70.log 2>&1; rc=$?; echo "ab small70 rc=$rc"; tail -4 $O/r2l_ab_small70.log
[ $rc -eq 0 ] || exit 1
timeout 150 python tools/ab_knobs.py 60x52x48 "$CFG" 1 > $O/r2l_ab_small.log 2>&1; rc=$?; echo "ab small rc=$rc"; tail -4 $O/r2l_ab_small.log
[ $rc -eq 0 ] || exit 1
timeout 500 python tools/ab_knobs.py 1536x1204x70 "$CFG" 5 > $O/r2l_ab_mid.log 2>&1; echo "ab mid rc=$?"; tail -4 $O/r2l_ab_mid.log
timeout 300 python tools/ab_knobs.py 400x317x48 "$CFG" 6 > $O/r2l_ab_core2.log 2>&1; echo "ab core2 rc=$?"; tail -4 $O/r2l_ab_core2.log
timeout 200 python tools/trace_pipeline.py 1536x1204x70 A "WT_CONV=-1" > $O/r2l_trace_A_fold.log 2>&1; echo "trace rc=$?"; tail -15 $O/r2l_trace_A_fold.log
