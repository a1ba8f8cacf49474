#!/bin/bash
# round 2, call I (1 GPU): full GPU test suite + the N=1 bench line of the final bench.py + GPU reference table
set -u
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r2i_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r2i_pytest.log
timeout 200 python __graft_entry__.py smoke > $O/r2i_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/r2i_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > $O/r2i_bench_ng5.json 2> $O/r2i_bench_ng5.err; echo "bench rc=$?"; tail -12 $O/r2i_bench_ng5.err
timeout 600 python tools/gpu_reference_bench.py core2 > $O/r2i_gpu_reference_core2.md 2> $O/r2i_gpu_reference_core2.err; echo "gpu ref core2 rc=$?"; cat $O/r2i_gpu_reference_core2.md
python - <<'PY'
import json
a=json.loads(open('gpurun_out/r2i_bench_ng5.json').read())
print("ms/step", a['ms_per_step'], "e2e", a['e2e']['ms_per_step'], "single", a['e2e']['single_tracer']['ms_per_step'], "packed", (a['e2e'].get('packed_host') or {}).get('ms_per_step'))
print("gpu_reference", a.get('gpu_reference'))
print("roofline", a['roofline']['kernels'])
PY
