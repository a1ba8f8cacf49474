#!/bin/bash
# round 2, call M (1 GPU): a1 folded (new default) x warp counts x ring depth
set -u
O=gpurun_out
mkdir -p $O
CFG="WT_CONV=2;WT_CONV=-1;WT_CONV=-1,WT_WARPS_A=24,WT_ISSUERS=3;WT_CONV=-1,WT_WARPS_A=23,WT_ISSUERS=4;WT_CONV=-1,WT_STAGES=3;WT_CONV=-1,WT_STAGES=3,WT_ISSUERS=4"
timeout 150 python tools/ab_knobs.py 96x74x70 "$CFG" 1 > $O/r2m_ab_small70.log 2>&1; rc=$?; echo "ab small70 rc=$rc"; tail -6 $O/r2m_ab_small70.log
[ $rc -eq 0 ] || exit 1
timeout 500 python tools/ab_knobs.py 1536x1204x70 "$CFG" 5 > $O/r2m_ab_mid.log 2>&1; echo "ab mid rc=$?"; tail -6 $O/r2m_ab_mid.log
CFG2="WT_CONV=2;WT_CONV=-1;WT_CONV=-1,WT_WARPS_A=24,WT_ISSUERS=3;WT_CONV=-1,WT_ISSUERS=4;WT_CONV=-1,WT_STAGES=2;WT_CONV=-1,WT_STAGES=4,WT_SMEM=56000"
timeout 300 python tools/ab_knobs.py 400x317x48 "$CFG2" 6 > $O/r2m_ab_core2.log 2>&1; echo "ab core2 rc=$?"; tail -6 $O/r2m_ab_core2.log
timeout 300 python tools/ab_knobs.py 2048x1560x80 "WT_CONV=2;WT_CONV=-1" 3 > $O/r2m_ab_dart.log 2>&1; echo "ab dart rc=$?"; tail -2 $O/r2m_ab_dart.log
