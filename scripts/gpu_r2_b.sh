#!/bin/bash
# round 2: new lane kernels -- parity, then the small configurations
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_fullsize.py::test_fixed_point_family_at_the_benchmark_shape > gpurun_out/pytest_b.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_b.log
for W in cfg2 cfg5 cfg4; do
  timeout 300 python bench.py --workload $W --steps 10 --warmup 3 --no-wrap > gpurun_out/bench_${W}_b.json 2> gpurun_out/bench_${W}_b.err; echo "bench $W rc=$?"
  python -c "
import json; d=json.loads(open('gpurun_out/bench_${W}_b.json').read().strip().splitlines()[-1]); print('$W', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), {k: round(v.get('ms_per_launch_per_restart', v.get('ms_per_launch', 0)), 4) for k, v in d['kernels'].items()}, d.get('fit_call'))"
done
VIREO_B200_ROWS_LANE=1 timeout 300 python bench.py --workload cfg2 --steps 10 --warmup 3 --no-wrap --no-cpu --no-fixed32 > gpurun_out/bench_cfg2_oldrule.json 2>/dev/null
python -c "
import json; d=json.loads(open('gpurun_out/bench_cfg2_oldrule.json').read().strip().splitlines()[-1]); print('cfg2 old rule', round(d['value'],1), {k: round(v.get('ms_per_launch_per_restart', 0), 4) for k, v in d['kernels'].items()})"
