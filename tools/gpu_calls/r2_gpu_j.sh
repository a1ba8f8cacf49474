#!/bin/bash
# round 2, call J (1 GPU): ncu evidence of the shipping build.  Every ncu pass runs after the same command
# exited 0 without ncu; numbers printed under ncu are never bench values.
set -u
O=gpurun_out
mkdir -p $O
B="python bench.py --steps 2 --warmup 3 --no-refseq --no-cpu --no-gpu-reference --no-parity --preroll 0 --e2e-tracers 1 --e2e-steps 1"
timeout 400 $B > $O/r2j_bench_plain_for_launch_list.json 2> $O/r2j_bench_plain.err; echo "plain rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2j_launches_bench_ng5.csv $B > $O/r2j_ncu_launches.log 2>&1; echo "launch list rc=$?"
# DRAM bytes per launch of the two fused kernels on the NG5 workload (single pass per metric: no 60 GB save/restore)
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_phase_warp -c 8 --csv --log-file $O/r2j_ng5_dram_bytes_launches.csv $B > $O/r2j_ncu_dram.log 2>&1; echo "dram bytes rc=$?"
# --set full of the two fused kernels on the 1.8M-node nl=70 mesh (packed), after the plain run
timeout 300 python tools/ncu_tile.py 1536x1204x70 phaseA_warp,phaseB_warp 3 1 packed > $O/r2j_plain_mid.log 2>&1 && \
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_phase_warp --launch-skip 2 -c 2 -o $O/prof_r2_mid -f \
    python tools/ncu_tile.py 1536x1204x70 phaseA_warp,phaseB_warp 2 1 packed > $O/r2j_ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la $O/prof_r2_mid.ncu-rep
