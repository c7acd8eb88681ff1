#!/bin/bash
# env-var sweep of one bench workload on the GPU box; prints kernel times per setting
mkdir -p gpurun_out
for spec in "$@"; do
  echo "=== $spec"
  timeout 300 env $spec python bench.py --workload ${WL:-cfg3} --steps 2 --warmup 2 --no-cpu 2>&1 | python -c "
import sys, json
for line in sys.stdin:
    line=line.strip()
    if not line.startswith('{'): print(line[:300]); continue
    d=json.loads(line)
    k=d['kernels']
    print('it/s %.1f  ms/iter %.3f  e2e %.1f | snp %.4f cell %.4f gt %.4f' % (d['value'], d['ms_per_iteration_per_restart'], d['e2e']['value'], k['k_snp']['ms_per_launch_per_restart'], k['k_cell']['ms_per_launch_per_restart'], k['k_gt']['ms_per_launch_per_restart']))
"
done
