#!/bin/bash
# vireo_wrap with the host draws overlapped with the warm-up fits: wrapper goldens, then the whole-call timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "wrap or cfg4 or doublet" > gpurun_out/pytest_n.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_n.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --no-fixed32 > gpurun_out/bench_cfg3_n.json 2> gpurun_out/bench_cfg3_n.err; echo "bench rc=$?"; tail -c 300 gpurun_out/bench_cfg3_n.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_cfg3_n.json').read().strip().splitlines()[-1]); print(round(d['value'],1), round(d['e2e']['value'],1)); print(d['wrap'])"
