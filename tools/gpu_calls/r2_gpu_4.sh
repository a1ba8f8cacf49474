#!/bin/bash
# round 2, 4 GPUs of one box: NG5 and DART at N = 4 and N = 2 with the final build (parity gate, digest, halo figures)
set -u
O=gpurun_out
mkdir -p $O
for W in ng5 dart; do
  for N in 4 2; do
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$N bench.py --gpus $N --steps 20 --warmup 5 --workload $W --no-refseq > $O/r2_4_bench_${W}_${N}gpu.json 2> $O/r2_4_bench_${W}_${N}gpu.err; echo "bench $W N=$N rc=$?"
  done
done
python - <<'PY'
import json
for w in ("ng5","dart"):
    for n in (4,2):
        try:
            a=json.loads(open(f'gpurun_out/r2_4_bench_{w}_{n}gpu.json').read())
            print(w, n, "ms/step", round(a['ms_per_step'],3), "G/s", round(a['value']/1e9,1), "frac", round(a['hbm']['frac_of_peak'],3), "parity", a['parity']['ok'], "digest", a['digest']['fct_plus'], "halo", {k:round(v,3) for k,v in a['halo'].items() if k!='how'}, "e2e", round(a['e2e']['ms_per_step'],1), round(a['e2e']['packed_host']['ms_per_step'],1), a['clocks']['sm_mhz'])
        except Exception as e: print(w, n, "failed", e)
PY
