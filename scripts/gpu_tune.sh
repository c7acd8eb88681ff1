#!/bin/bash
# quick tuning sweep on the GPU box: env-var knobs x cfg3 bench (no CPU leg)
mkdir -p gpurun_out
for spec in "$@"; do
  echo "=== $spec"
  env $spec python bench.py --workload ${WL:-cfg3} --steps 2 --warmup 2 --no-cpu 2>&1 | python -c "
import sys, json
for line in sys.stdin:
    line=line.strip()
    if not line.startswith('{'): print(line[:300]); continue
    d=json.loads(line)
    print('it/s %.1f  ms/iter %.3f  e2e %.1f' % (d['value'], d['ms_per_iteration_per_restart'], d['e2e']['value']))
    for k,v in d['kernels'].items(): print('   %-12s %.4f ms share %.3f' % (k, v['ms_per_launch_per_restart'], v['share']))
    print('   elbo', d['elbo_final'])
"
done
