#!/bin/bash
# One GPU box, one call: the round's measured evidence (gpurun_out/ is scratch; the summaries are
# copied into profiles/ afterwards).  Every ncu pass runs after the same command exited 0 without ncu.
set -u
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/final_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $O/final_pytest.log
timeout 200 python __graft_entry__.py smoke > $O/final_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/final_smoke.log
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > $O/final_bench_reference.json 2> $O/final_bench_reference.err; echo "reference arm rc=$?"
timeout 500 python bench.py --steps 10 --warmup 3 > $O/final_bench_ng5.json 2> $O/final_bench_ng5.err; echo "bench rc=$?"; tail -3 $O/final_bench_ng5.err
timeout 300 python bench.py --workload core2 --steps 20 --warmup 5 --no-refseq --no-cpu > $O/final_bench_core2.json 2> $O/final_bench_core2.err; echo "bench core2 rc=$?"
# launch list of the bench command (per-launch times are cold-cache and serialised: shares, not absolutes)
timeout 300 python bench.py --steps 2 --warmup 3 --no-refseq --no-cpu > $O/final_bench_plain_for_launch_list.json 2> /dev/null && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/final_launches_bench_ng5.csv \
    python bench.py --steps 2 --warmup 3 --no-refseq --no-cpu > $O/final_ncu_launches.log 2>&1; echo "launch list rc=$?"
# --set full of the two fused kernels on the 1.8M-node nl=70 mesh (packed), after the plain run
timeout 300 python tools/ncu_tile.py 1536x1204x70 phaseA_warp,phaseB_warp 3 1 packed > $O/final_plain_mid.log 2>&1 && \
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_phase_warp --launch-skip 2 -c 2 -o $O/prof_final_mid -f \
    python tools/ncu_tile.py 1536x1204x70 phaseA_warp,phaseB_warp 2 1 packed > $O/final_ncu_full.log 2>&1; echo "ncu full rc=$?"
