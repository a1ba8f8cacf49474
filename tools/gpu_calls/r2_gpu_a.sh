#!/bin/bash
# round 2, call A (1 GPU): GPU test suite, N=1 bench line (parity gate + digest inside), the same-mesh CPU arm
set -u
O=gpurun_out
mkdir -p $O
nproc > $O/r2a_host.txt; free -g >> $O/r2a_host.txt; lscpu | grep -i -E "numa|model name|socket" >> $O/r2a_host.txt; nvidia-smi topo -m >> $O/r2a_host.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r2a_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r2a_pytest.log
timeout 200 python __graft_entry__.py smoke > $O/r2a_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/r2a_smoke.log
timeout 800 python bench.py --steps 20 --warmup 5 > $O/r2a_bench_ng5.json 2> $O/r2a_bench_ng5.err; echo "bench rc=$?"; tail -4 $O/r2a_bench_ng5.err
timeout 800 python bench.py --impl reference --steps 3 --warmup 1 > $O/r2a_bench_reference.json 2> $O/r2a_bench_reference.err; echo "reference arm rc=$?"; tail -2 $O/r2a_bench_reference.err
python - <<'PY'
import json
a=json.loads(open('gpurun_out/r2a_bench_ng5.json').read()); b=json.loads(open('gpurun_out/r2a_bench_reference.json').read())
print("gpu digest", a.get('digest')); print("cpu digest", b.get('digest')); print("equal:", a.get('digest')==b.get('digest'))
print("parity", a.get('parity',{}).get('ok'), "ms/step", a.get('ms_per_step'), "roofline", a['roofline']['kernels'], "e2e ms", a['e2e']['ms_per_step'])
print("cpu arm", b['value'], b['cpu_baseline']['cores'], b['ms_per_step'])
PY
