#!/bin/bash
# round 2: segment format v4 (32-bit records, ring rows, multi-window look-ahead) -- format check, parity, bench, ring geometry sweep
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fullsize.py -m gpu -q -x -k "segment_format" > gpurun_out/pytest_e0.log 2>&1; rc=$?; echo "format rc=$rc"; tail -5 gpurun_out/pytest_e0.log
if [ $rc -ne 0 ]; then grep -n "Error\|assert" gpurun_out/pytest_e0.log | head -20; exit 1; fi
timeout 1500 python -m pytest tests -m gpu -q --maxfail=8 > gpurun_out/pytest_e.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_e.log
show() { python -c "
import json,sys; d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[2], round(d['value'],1), 'e2e', round(d['e2e']['value'],1), {k: round(v.get('ms_per_launch_per_restart', v.get('ms_per_launch', 0)), 4) for k, v in d['kernels'].items()}, d.get('parity_check'))" $1 "$2"; }
timeout 600 python bench.py --steps 5 --warmup 3 --no-wrap > gpurun_out/bench_cfg3_e.json 2> gpurun_out/bench_cfg3_e.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_cfg3_e.err
show gpurun_out/bench_cfg3_e.json "cfg3 default"
for G in "384 4 3" "192 8 6" "512 3 2" "256 6 3" "128 12 9"; do
  set -- $G
  VIREO_B200_SEG_WR64=$1 VIREO_B200_SEG_NB64=$2 VIREO_B200_SEG_SPAN64=$3 timeout 300 python bench.py --steps 5 --warmup 3 --no-wrap --no-cpu --no-fixed32 > gpurun_out/bench_cfg3_e_$1_$2_$3.json 2>/dev/null
  show gpurun_out/bench_cfg3_e_$1_$2_$3.json "cfg3 win=$1 nb=$2 span=$3"
done
timeout 300 python bench.py --workload cfg4 --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_cfg4_e.json 2>/dev/null; show gpurun_out/bench_cfg4_e.json cfg4
