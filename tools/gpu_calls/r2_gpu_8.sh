#!/bin/bash
# round 2, 8 GPUs of one box: NCCL parity test, NG5 and DART bench lines (parity gate + digest inside), host-link ceiling
set -u
O=gpurun_out
mkdir -p $O
nproc > $O/r2_8_host.txt; free -g >> $O/r2_8_host.txt; lscpu | grep -i -E "numa|model name|socket" >> $O/r2_8_host.txt; nvidia-smi topo -m >> $O/r2_8_host.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -rA > $O/r2_8_pytest_multi.log 2>&1; echo "pytest multi rc=$?"; tail -5 $O/r2_8_pytest_multi.log
timeout 600 $TR --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 > $O/r2_8_bench_ng5.json 2> $O/r2_8_bench_ng5.err; echo "bench ng5 N=8 rc=$?"; tail -3 $O/r2_8_bench_ng5.err
timeout 600 $TR --master-port 29522 bench.py --gpus 8 --steps 20 --warmup 5 --workload dart --no-refseq > $O/r2_8_bench_dart.json 2> $O/r2_8_bench_dart.err; echo "bench dart N=8 rc=$?"; tail -3 $O/r2_8_bench_dart.err
timeout 300 $TR --master-port 29523 tools/h2d_concurrent.py > $O/r2_8_h2d_concurrent.json 2> $O/r2_8_h2d_concurrent.err; echo "h2d rc=$?"; cat $O/r2_8_h2d_concurrent.json | cut -c1-600
python - <<'PY'
import json
for n in ("ng5","dart"):
    try:
        a=json.loads(open(f'gpurun_out/r2_8_bench_{n}.json').read())
        print(n, "ms/step", a.get('ms_per_step'), "parity", a.get('parity',{}).get('ok'), "digest", a.get('digest'), "halo", {k:v for k,v in (a.get('halo') or {}).items() if k!='how'}, "e2e", a['e2e']['ms_per_step'], (a['e2e'].get('packed_host') or {}).get('ms_per_step'))
    except Exception as e: print(n, "failed", e)
PY
