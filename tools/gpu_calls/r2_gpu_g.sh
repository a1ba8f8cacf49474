#!/bin/bash
# round 2, call G (1 GPU): how the producer warps wait (polling every ~26 ns = 15 % of the issued instructions in round 1)
set -u
O=gpurun_out
mkdir -p $O
CFG="WT_OPT=6;WT_OPT=22;WT_OPT=38;WT_OPT=70;WT_OPT=30;WT_OPT=46;WT_OPT=78;WT_OPT=7"
timeout 200 python tools/ab_knobs.py 96x74x70 "$CFG" 1 > $O/r2g_ab_small70.log 2>&1; rc=$?; echo "ab small70 rc=$rc"; tail -8 $O/r2g_ab_small70.log
[ $rc -eq 0 ] || exit 1
timeout 600 python tools/ab_knobs.py 1536x1204x70 "$CFG" 5 > $O/r2g_ab_mid.log 2>&1; echo "ab mid rc=$?"; tail -8 $O/r2g_ab_mid.log
timeout 200 python tools/trace_pipeline.py 1536x1204x70 A "WT_OPT=22" > $O/r2g_trace_A_opt22.log 2>&1; echo "trace rc=$?"; tail -15 $O/r2g_trace_A_opt22.log
timeout 200 python tools/trace_pipeline.py 1536x1204x70 A "WT_OPT=38" > $O/r2g_trace_A_opt38.log 2>&1; echo "trace rc=$?"; tail -15 $O/r2g_trace_A_opt38.log
