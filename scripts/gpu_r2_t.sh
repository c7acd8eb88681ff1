#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/diag_t.jsonl
for V in "" _lane5 _lane6; do
  VIREO_B200_LIB=$PWD/vireo_b200/libvireo_b200${V}.so timeout 300 python bench.py --workload cfg2 --steps 10 --warmup 3 --no-cpu --no-wrap --no-fixed32 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('cfg2 $V', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), {k: round(v.get('ms_per_launch_per_restart',0),4) for k,v in d['kernels'].items()})"
  VIREO_B200_LIB=$PWD/vireo_b200/libvireo_b200${V}.so timeout 300 python bench.py --workload cfg5 --steps 10 --warmup 3 --no-cpu 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('cfg5 $V', round(d['value'],1), d['fit_call']['s'], {k: round(v.get('ms_per_launch',0),4) for k,v in d['kernels'].items()})"
done
