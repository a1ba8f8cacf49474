#!/bin/bash
# round 2, call Y (1 GPU): final build, the other single-GPU BASELINE.json configurations
set -u
O=gpurun_out
mkdir -p $O
timeout 400 python bench.py --workload dart --steps 20 --warmup 5 --no-refseq --no-cpu --no-gpu-reference > $O/r2y_bench_dart.json 2> $O/r2y_bench_dart.err; echo "bench dart rc=$?"
timeout 200 python bench.py --workload core2 --steps 50 --warmup 10 --no-refseq --no-cpu --no-gpu-reference > $O/r2y_bench_core2.json 2> $O/r2y_bench_core2.err; echo "bench core2 rc=$?"
timeout 200 python tools/multi_tracer_bench.py core2 12 50 > $O/r2y_multi_tracer_core2x12.log 2>&1; echo "multi tracer rc=$?"; tail -6 $O/r2y_multi_tracer_core2x12.log
python - <<'PY'
import json
for n in ("dart","core2"):
    a=json.loads(open(f'gpurun_out/r2y_bench_{n}.json').read())
    print(n, "ms/step", a['ms_per_step'], "G/s", a['value']/1e9, "hbm", a['hbm']['frac_of_peak'], "kern", {k:round(v['frac'],3) for k,v in a['roofline']['kernels'].items()}, a['parity']['ok'], a['clocks'])
PY
