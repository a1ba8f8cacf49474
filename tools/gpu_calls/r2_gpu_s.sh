#!/bin/bash
# round 2, call S (1 GPU): final validation -- GPU test suite, smoke, the default bench line
set -u
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r2s_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r2s_pytest.log
timeout 200 python __graft_entry__.py smoke > $O/r2s_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/r2s_smoke.log
timeout 900 python bench.py > $O/r2s_bench_default.json 2> $O/r2s_bench_default.err; echo "bench rc=$?"; tail -3 $O/r2s_bench_default.err
python - <<'PY'
import json
a=json.loads(open('gpurun_out/r2s_bench_default.json').read())
print("ms/step", a['ms_per_step'], "clocks", a['clocks'], "parity", a['parity']['ok'], "digest", a['digest']['fct_plus'], "roofline", a['roofline']['frac'], a['roofline']['traffic'], "e2e", a['e2e']['ms_per_step'], "launches", a['gpu_launches'], "setup", a['setup_s'])
PY
