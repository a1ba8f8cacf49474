#!/bin/bash
# round 2, call C (1 GPU): interleaved A/B of the register re-allocation variants (setmaxnreg) against the round-1 shape
set -u
O=gpurun_out
mkdir -p $O
# hang guard first: a wrong register budget makes setmaxnreg.inc wait forever
timeout 150 python tools/ab_knobs.py 60x52x48 "WT_REGS=0;WT_REGS=80;WT_REGS=88" 1 > $O/r2c_ab_small.log 2>&1; rc=$?; echo "ab small rc=$rc"; tail -3 $O/r2c_ab_small.log
[ $rc -eq 0 ] || exit 1
timeout 150 python tools/ab_knobs.py 96x74x70 "WT_REGS=0;WT_REGS=80;WT_REGS=88" 1 > $O/r2c_ab_small70.log 2>&1; rc=$?; echo "ab small70 rc=$rc"; tail -3 $O/r2c_ab_small70.log
[ $rc -eq 0 ] || exit 1
timeout 420 python tools/ab_knobs.py 1536x1204x70 "WT_REGS=0;WT_REGS=80;WT_REGS=88;WT_REGS=80,WT_OPT=6;WT_REGS=80,WT_OPT=3" 4 > $O/r2c_ab_mid.log 2>&1; echo "ab mid rc=$?"; tail -6 $O/r2c_ab_mid.log
timeout 300 python tools/ab_knobs.py 400x317x48 "WT_REGS=0;WT_REGS=80;WT_REGS=88;WT_REGS=80,WT_OPT=2" 6 > $O/r2c_ab_core2.log 2>&1; echo "ab core2 rc=$?"; tail -5 $O/r2c_ab_core2.log
