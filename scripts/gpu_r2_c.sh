#!/bin/bash
# round 2: fused tails + lane-kernel variants -- parity, then the small configurations and an ncu capture of cfg2
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_c.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_c.log
show() { python -c "
import json,sys; d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[2], round(d['value'],1), 'e2e', round(d['e2e']['value'],1), {k: round(v.get('ms_per_launch_per_restart', v.get('ms_per_launch', 0)), 4) for k, v in d['kernels'].items()}, d.get('fit_call', {}).get('s'))" $1 "$2"; }
for DEEP in 1 0; do for FUSE in 1 0; do
  VIREO_B200_LANE_DEEP=$DEEP VIREO_B200_FUSE=$FUSE timeout 300 python bench.py --workload cfg2 --steps 10 --warmup 3 --no-wrap --no-cpu --no-fixed32 > gpurun_out/bench_cfg2_d${DEEP}f${FUSE}.json 2>/dev/null
  show gpurun_out/bench_cfg2_d${DEEP}f${FUSE}.json "cfg2 deep=$DEEP fuse=$FUSE"
done; done
for DEEP in 1 0; do
  VIREO_B200_LANE_DEEP=$DEEP timeout 300 python bench.py --workload cfg5 --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_cfg5_d${DEEP}.json 2>/dev/null
  show gpurun_out/bench_cfg5_d${DEEP}.json "cfg5 deep=$DEEP"
done
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-fixed32 > gpurun_out/bench_cfg3_c.json 2> gpurun_out/bench_cfg3_c.err
show gpurun_out/bench_cfg3_c.json "cfg3"
python -c "
import json; d=json.loads(open('gpurun_out/bench_cfg3_c.json').read().strip().splitlines()[-1]); print(json.dumps(d['wrap'])); print(json.dumps(d['doublet_ms']))"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_(cell|snp)_lane' -s 8 -c 4 -f -o gpurun_out/prof_cfg2_lane python bench.py --workload cfg2 --steps 1 --warmup 1 --no-cpu --no-wrap --no-fixed32 > gpurun_out/ncu_cfg2_lane.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/prof_cfg2_lane.ncu-rep --page raw --csv > gpurun_out/prof_cfg2_lane_raw.csv 2>/dev/null
