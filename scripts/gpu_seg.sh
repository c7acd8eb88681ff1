#!/bin/bash
# first contact of the window-segment kernels with the GPU: small parity tests, then cfg3 timings per family
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "seg and not full_size" --durations=5 > gpurun_out/pytest_seg.log 2>&1; echo "pytest seg rc=$?"
tail -25 gpurun_out/pytest_seg.log
for spec in "$@"; do
  echo "=== $spec"
  timeout 600 env $spec python bench.py --workload ${WL:-cfg3} --steps 2 --warmup 2 --no-cpu 2>&1 | python -c "
import sys, json
for line in sys.stdin:
    line=line.strip()
    if not line.startswith('{'): print(line[:300]); continue
    d=json.loads(line)
    print('it/s %.1f  ms/iter %.3f  e2e %.1f' % (d['value'], d['ms_per_iteration_per_restart'], d['e2e']['value']))
    for k,v in d['kernels'].items(): print('   %-12s %.4f ms share %.3f' % (k, v['ms_per_launch_per_restart'], v['share']))
    print('   elbo', d['elbo_final'], d.get('parity_check'))
"
done
