#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_cfg3_u.json 2> gpurun_out/bench_cfg3_u.err; echo "bench cfg3 rc=$?"; tail -c 200 gpurun_out/bench_cfg3_u.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_cfg3_u.json').read().strip().splitlines()[-1])
print('cfg3', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), {k: round(v.get('ms_per_launch_per_restart', 0), 4) for k, v in d['kernels'].items()}, 'fixed32', round(d['fixed32']['value'],1), 'wrap', round(d['wrap']['total_s'],3), 'xbar', round(d['roofline']['crossbar']['frac'],3), 'frac', round(d['roofline']['frac'],4), 'traffic', d['roofline']['traffic'])"
for W in cfg2 cfg4 cfg5; do
  timeout 300 python bench.py --workload $W --steps 10 --warmup 3 > gpurun_out/bench_${W}_u.json 2> gpurun_out/bench_${W}_u.err; echo "bench $W rc=$?"
  python -c "
import json; d=json.loads(open('gpurun_out/bench_${W}_u.json').read().strip().splitlines()[-1]); print('$W', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d.get('fit_call',{}).get('s'))"
done
