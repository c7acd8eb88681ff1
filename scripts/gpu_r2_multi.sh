#!/bin/bash
# round 2, N GPUs (N = $1): NCCL leg parity of the sharded fit / doublet / wrapper, then the bench line at N
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR scripts/gpu_sharded_parity.py > gpurun_out/sharded_parity_${N}gpu.json 2> gpurun_out/sharded_parity_${N}gpu.err; echo "parity rc=$?"
tail -c 1500 gpurun_out/sharded_parity_${N}gpu.json; tail -5 gpurun_out/sharded_parity_${N}gpu.err
timeout 900 $TR bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_cfg3_${N}gpu.json 2> gpurun_out/bench_cfg3_${N}gpu.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/bench_cfg3_${N}gpu.json; tail -5 gpurun_out/bench_cfg3_${N}gpu.err
