#!/bin/bash
# round 2, call E (1 GPU): pipeline v2 (consumers convert, fast hand-over, 24 consumers) against the round-1 shape
set -u
O=gpurun_out
mkdir -p $O
CFG="WT_CONV=2;WT_CONV=0;WT_CONV=0,WT_OPT=14;WT_CONV=0,WT_OPT=14,WT_WARPS_A=24,WT_WARPS_B=24,WT_ISSUERS=3;WT_CONV=0,WT_OPT=14,WT_REGS=80;WT_CONV=2,WT_OPT=14"
timeout 150 python tools/ab_knobs.py 96x74x70 "$CFG" 1 > $O/r2e_ab_small70.log 2>&1; rc=$?; echo "ab small70 rc=$rc"; tail -7 $O/r2e_ab_small70.log
[ $rc -eq 0 ] || exit 1
timeout 150 python tools/ab_knobs.py 60x52x48 "$CFG" 1 > $O/r2e_ab_small.log 2>&1; rc=$?; echo "ab small rc=$rc"; tail -7 $O/r2e_ab_small.log
[ $rc -eq 0 ] || exit 1
timeout 420 python tools/ab_knobs.py 1536x1204x70 "$CFG" 4 > $O/r2e_ab_mid.log 2>&1; echo "ab mid rc=$?"; tail -7 $O/r2e_ab_mid.log
timeout 300 python tools/ab_knobs.py 400x317x48 "$CFG" 6 > $O/r2e_ab_core2.log 2>&1; echo "ab core2 rc=$?"; tail -7 $O/r2e_ab_core2.log
timeout 200 python tools/trace_pipeline.py 1536x1204x70 A "WT_CONV=0,WT_OPT=14,WT_WARPS_A=24,WT_WARPS_B=24,WT_ISSUERS=3" > $O/r2e_trace_A_v2.log 2>&1; echo "trace A rc=$?"; tail -16 $O/r2e_trace_A_v2.log
timeout 200 python tools/trace_pipeline.py 1536x1204x70 B "WT_CONV=0,WT_OPT=14,WT_WARPS_A=24,WT_WARPS_B=24,WT_ISSUERS=3" > $O/r2e_trace_B_v2.log 2>&1; echo "trace B rc=$?"; tail -12 $O/r2e_trace_B_v2.log
