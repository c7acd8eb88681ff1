#!/bin/bash
# row-split cell pass: the new test, the cfg4 full-size test, the segment parity subset, then timings
mkdir -p gpurun_out; rm -f gpurun_out/diag_m.jsonl
timeout 900 python -m pytest tests/test_gpu_fullsize.py -m gpu -q -x -k "row_split or cfg4 or format" > gpurun_out/pytest_m0.log 2>&1; rc=$?; echo "split tests rc=$rc"; tail -4 gpurun_out/pytest_m0.log
if [ $rc -ne 0 ]; then grep -n "Error\|assert\|mismatch" gpurun_out/pytest_m0.log | head -20; exit 1; fi
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "seg" > gpurun_out/pytest_m1.log 2>&1; echo "parity rc=$?"; tail -3 gpurun_out/pytest_m1.log
for S in 1 0; do
  VIREO_B200_SEG_SPLIT=$S timeout 300 python scripts/time_passes.py cfg4 10 "cfg4 split=$S" 2>&1 | tail -1 | tee -a gpurun_out/diag_m.jsonl | cut -c1-200
  for N in 2 4 8; do VIREO_B200_SEG_SPLIT=$S timeout 300 python scripts/time_shard.py $N 10 2>&1 | tail -1 | tee -a gpurun_out/diag_m.jsonl | cut -c1-200; done
done
