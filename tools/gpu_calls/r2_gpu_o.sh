#!/bin/bash
# round 2, call O (1 GPU): final build -- bench lines of every single-GPU BASELINE.json configuration
set -u
O=gpurun_out
mkdir -p $O
timeout 200 python __graft_entry__.py smoke > $O/r2o_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/r2o_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > $O/r2o_bench_ng5.json 2> $O/r2o_bench_ng5.err; echo "bench ng5 rc=$?"; tail -4 $O/r2o_bench_ng5.err
timeout 600 python bench.py --workload dart --steps 20 --warmup 5 --no-refseq --no-cpu > $O/r2o_bench_dart.json 2> $O/r2o_bench_dart.err; echo "bench dart rc=$?"
timeout 300 python bench.py --workload core2 --steps 50 --warmup 10 --no-refseq --no-cpu > $O/r2o_bench_core2.json 2> $O/r2o_bench_core2.err; echo "bench core2 rc=$?"
timeout 300 python tools/multi_tracer_bench.py core2 12 50 > $O/r2o_multi_tracer_core2x12.log 2>&1; echo "multi tracer rc=$?"; tail -6 $O/r2o_multi_tracer_core2x12.log
python - <<'PY'
import json
for n in ("ng5","dart","core2"):
    a=json.loads(open(f'gpurun_out/r2o_bench_{n}.json').read())
    print(n, "ms/step", a['ms_per_step'], "G/s", a['value']/1e9, "hbm", a['hbm']['frac_of_peak'], "kern", {k:round(v['frac'],3) for k,v in a['roofline']['kernels'].items()}, "e2e", a['e2e']['ms_per_step'], (a['e2e'].get('packed_host') or {}).get('ms_per_step'), "gpuref", (a.get('gpu_reference') or {}).get('step_ms'), a['parity']['ok'], a['clocks'])
PY
