#!/bin/bash
# round 2: after the shared-memory opt-in fix -- whole GPU suite, headline bench, small configurations, launch list
mkdir -p gpurun_out
timeout 600 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -q --maxfail=12 --durations=15 > gpurun_out/pytest_d.log 2>&1; echo "pytest rc=$?"
tail -30 gpurun_out/pytest_d.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_cfg3_d.json 2> gpurun_out/bench_cfg3_d.err; echo "bench cfg3 rc=$?"
tail -c 400 gpurun_out/bench_cfg3_d.err
head -c 6000 gpurun_out/bench_cfg3_d.json
for W in cfg2 cfg4 cfg5; do
  timeout 300 python bench.py --workload $W --steps 10 --warmup 3 > gpurun_out/bench_${W}_d.json 2> gpurun_out/bench_${W}_d.err; echo "bench $W rc=$?"
done
