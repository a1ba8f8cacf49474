#!/bin/bash
# round 2, call U (1 GPU): ncu --set full of the FINAL fused kernels on the 1.8M-node nl=70 mesh (after the plain run)
set -u
O=gpurun_out
mkdir -p $O
timeout 300 python tools/ncu_tile.py 1536x1204x70 phaseA_warp,phaseB_warp 3 1 packed > $O/r2u_plain_mid.log 2>&1 && \
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_phase_warp --launch-skip 2 -c 2 -o $O/prof_r2_final_mid -f \
    python tools/ncu_tile.py 1536x1204x70 phaseA_warp,phaseB_warp 2 1 packed > $O/r2u_ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la $O/prof_r2_final_mid.ncu-rep
