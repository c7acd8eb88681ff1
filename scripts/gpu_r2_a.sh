#!/bin/bash
# round 2, first GPU call: smoke, the whole GPU test-suite, the headline bench and the small configurations
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > gpurun_out/gpu.txt 2>&1
timeout 600 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
timeout 1500 python -m pytest tests -m gpu -q --maxfail=12 --durations=15 > gpurun_out/pytest_all.log 2>&1; echo "pytest rc=$?"
tail -30 gpurun_out/pytest_all.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_cfg3.json 2> gpurun_out/bench_cfg3.err; echo "bench cfg3 rc=$?"
for W in cfg2 cfg4 cfg5; do
  timeout 300 python bench.py --workload $W --steps 10 --warmup 3 > gpurun_out/bench_$W.json 2> gpurun_out/bench_$W.err; echo "bench $W rc=$?"
done
tail -c 600 gpurun_out/bench_cfg3.err
head -c 1500 gpurun_out/bench_cfg3.json
