#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fullsize.py -m gpu -q -x -k "row_split" > gpurun_out/pytest_o.log 2>&1; echo "split test rc=$?"; tail -4 gpurun_out/pytest_o.log
bash scripts/gpu_r2_sanitize.sh 2>&1 | grep -E "rc=|SUMMARY|passed|failed" 
