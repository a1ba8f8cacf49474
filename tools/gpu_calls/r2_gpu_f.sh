#!/bin/bash
# round 2, call F (1 GPU): fast hand-over x ring depth x converter warps
set -u
O=gpurun_out
mkdir -p $O
CFG="WT_OPT=6;WT_OPT=14;WT_OPT=14,WT_CONV=4;WT_OPT=14,WT_CONV=4,WT_ISSUERS=2;WT_STAGES=3,WT_OPT=2;WT_STAGES=3,WT_OPT=10;WT_STAGES=3,WT_OPT=14;WT_STAGES=3,WT_OPT=10,WT_CONV=4;WT_STAGES=3,WT_OPT=10,WT_ISSUERS=4"
timeout 200 python tools/ab_knobs.py 96x74x70 "$CFG" 1 > $O/r2f_ab_small70.log 2>&1; rc=$?; echo "ab small70 rc=$rc"; tail -9 $O/r2f_ab_small70.log
[ $rc -eq 0 ] || exit 1
timeout 600 python tools/ab_knobs.py 1536x1204x70 "$CFG" 4 > $O/r2f_ab_mid.log 2>&1; echo "ab mid rc=$?"; tail -9 $O/r2f_ab_mid.log
CFG2="WT_OPT=2;WT_OPT=10;WT_OPT=10,WT_CONV=4;WT_OPT=14;WT_STAGES=2,WT_OPT=6;WT_STAGES=2,WT_OPT=14;WT_STAGES=4,WT_OPT=10,WT_SMEM=56000"
timeout 400 python tools/ab_knobs.py 400x317x48 "$CFG2" 6 > $O/r2f_ab_core2.log 2>&1; echo "ab core2 rc=$?"; tail -7 $O/r2f_ab_core2.log
