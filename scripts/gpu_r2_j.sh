#!/bin/bash
# round 2: validation of the final segment kernels -- smoke, whole GPU suite, bench lines of every configuration
mkdir -p gpurun_out
timeout 600 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -q --maxfail=12 --durations=8 > gpurun_out/pytest_j.log 2>&1; echo "pytest rc=$?"
tail -14 gpurun_out/pytest_j.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_cfg3_j.json 2> gpurun_out/bench_cfg3_j.err; echo "bench cfg3 rc=$?"
tail -c 300 gpurun_out/bench_cfg3_j.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_cfg3_j.json').read().strip().splitlines()[-1])
print('cfg3', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), {k: round(v.get('ms_per_launch_per_restart', 0), 4) for k, v in d['kernels'].items()})
print('roofline', json.dumps(d['roofline'])[:600]); print('fixed32', json.dumps(d['fixed32'])[:700]); print('doublet', d['doublet_ms']); print('wrap', d['wrap']); print('parity', d['parity_check'], 'cold', d['cold_e2e'])"
for W in cfg2 cfg4 cfg5; do
  timeout 300 python bench.py --workload $W --steps 10 --warmup 3 > gpurun_out/bench_${W}_j.json 2> gpurun_out/bench_${W}_j.err; echo "bench $W rc=$?"
  python -c "
import json; d=json.loads(open('gpurun_out/bench_${W}_j.json').read().strip().splitlines()[-1]); print('$W', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), {k: round(v.get('ms_per_launch_per_restart', v.get('ms_per_launch', 0)), 4) for k, v in d['kernels'].items()}, d.get('fit_call'))"
done
