#!/bin/bash
# round 2, call V (1 GPU): tile-level L2 prefetch of the consumers' direct loads by the issuer warps (WT_OPT bit 128)
set -u
O=gpurun_out
mkdir -p $O
timeout 150 python tools/ab_knobs.py 60x52x48 "WT_OPT=2;WT_OPT=130" 1 > $O/r2v_ab_small.log 2>&1; rc=$?; echo "ab small rc=$rc"; tail -2 $O/r2v_ab_small.log
[ $rc -eq 0 ] || exit 1
timeout 150 python tools/ab_knobs.py 96x74x70 "WT_OPT=6;WT_OPT=134" 1 > $O/r2v_ab_small70.log 2>&1; rc=$?; echo "ab small70 rc=$rc"; tail -2 $O/r2v_ab_small70.log
[ $rc -eq 0 ] || exit 1
timeout 400 python tools/ab_knobs.py 1536x1204x70 "WT_OPT=6;WT_OPT=134" 6 > $O/r2v_ab_mid.log 2>&1; echo "ab mid rc=$?"; tail -2 $O/r2v_ab_mid.log
timeout 300 python tools/ab_knobs.py 400x317x48 "WT_OPT=2;WT_OPT=130" 8 > $O/r2v_ab_core2.log 2>&1; echo "ab core2 rc=$?"; tail -2 $O/r2v_ab_core2.log
