#!/bin/bash
# round 2: compute-sanitizer evidence.  memcheck over every family; racecheck on the segment
# kernels with the default build (cp.async.bulk window fills) and with the plain-store variant of the same protocol.
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
python scripts/sanitize_run.py all > gpurun_out/san_plain_default.txt 2>&1; echo "plain default rc=$?"
VIREO_B200_LIB=$PWD/vireo_b200/libvireo_b200_plainfill.so python scripts/sanitize_run.py all > gpurun_out/san_plain_plainfill.txt 2>&1; echo "plain plainfill rc=$?"
cat gpurun_out/san_plain_default.txt gpurun_out/san_plain_plainfill.txt
timeout 900 $CS --tool memcheck --print-limit 20 python scripts/sanitize_run.py all small > gpurun_out/memcheck_all.log 2>&1; echo "memcheck rc=$?"
tail -5 gpurun_out/memcheck_all.log
VIREO_B200_LIB=$PWD/vireo_b200/libvireo_b200_plainfill.so timeout 900 $CS --tool racecheck --racecheck-report analysis --print-limit 20 python scripts/sanitize_run.py seg small > gpurun_out/racecheck_plainfill.log 2>&1; echo "racecheck plainfill rc=$?"
tail -5 gpurun_out/racecheck_plainfill.log
timeout 900 $CS --tool racecheck --racecheck-report analysis --print-limit 20 python scripts/sanitize_run.py seg small > gpurun_out/racecheck_default.log 2>&1; echo "racecheck default rc=$?"
tail -5 gpurun_out/racecheck_default.log
timeout 600 $CS --tool synccheck --print-limit 20 python scripts/sanitize_run.py seg small > gpurun_out/synccheck_seg.log 2>&1; echo "synccheck rc=$?"
tail -3 gpurun_out/synccheck_seg.log
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "sharded or shards" > gpurun_out/pytest_sharded.log 2>&1; echo "pytest sharded rc=$?"; tail -3 gpurun_out/pytest_sharded.log
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --no-fixed32 > gpurun_out/bench_cfg3_wrap.json 2> gpurun_out/bench_cfg3_wrap.err; echo "bench rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/bench_cfg3_wrap.json').read().strip().splitlines()[-1]); print(json.dumps(d['wrap'])); print(d['value'], d['e2e']['value'])"
