#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fullsize.py -m gpu -q -x -k "segment_format or heavy_tailed or row_split" > gpurun_out/pytest_x.log 2>&1; rc=$?; echo "owner-split tests rc=$rc"; tail -5 gpurun_out/pytest_x.log
if [ $rc -ne 0 ]; then grep -n "Error\|assert\|mismatch" gpurun_out/pytest_x.log | head -20; exit 1; fi
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "seg" > gpurun_out/pytest_x1.log 2>&1; echo "parity rc=$?"; tail -3 gpurun_out/pytest_x1.log
