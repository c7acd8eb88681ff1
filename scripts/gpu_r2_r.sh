#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=5 > gpurun_out/pytest_r.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_r.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_cfg3_r.json 2> gpurun_out/bench_cfg3_r.err; echo "bench cfg3 rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/bench_cfg3_r.json').read().strip().splitlines()[-1])
print('cfg3', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), {k: round(v.get('ms_per_launch_per_restart', 0), 4) for k, v in d['kernels'].items()}, 'fixed32', round(d['fixed32']['value'],1), 'wrap', round(d['wrap']['total_s'],3), 'xbar', round(d['roofline']['crossbar']['frac'],3), 'frac', round(d['roofline']['frac'],4))"
