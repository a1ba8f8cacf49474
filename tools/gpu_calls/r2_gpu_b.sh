#!/bin/bash
# round 2, call B (2 GPUs): the NCCL parity test (tests/nccl_worker.py) and the N=2 bench line
set -u
O=gpurun_out
mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -rA > $O/r2b_pytest_multi.log 2>&1; echo "pytest multi rc=$?"; tail -5 $O/r2b_pytest_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > $O/r2b_bench_ng5_2gpu.json 2> $O/r2b_bench_ng5_2gpu.err; echo "bench N=2 rc=$?"; tail -4 $O/r2b_bench_ng5_2gpu.err
python - <<'PY'
import json
a=json.loads(open('gpurun_out/r2b_bench_ng5_2gpu.json').read())
print("digest", a.get('digest')); print("parity", a.get('parity')); print("ms/step", a.get('ms_per_step'), "halo", a.get('halo'), "preroll", a.get('preroll_s'), "e2e", a['e2e']['ms_per_step'])
PY
