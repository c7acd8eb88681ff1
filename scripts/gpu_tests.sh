#!/bin/bash
# run on the GPU box via gpurun: GPU parity tests, output under gpurun_out/
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout ${1:-900} python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"
tail -60 gpurun_out/pytest.log
