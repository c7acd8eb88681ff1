#!/bin/bash
# 4-landing-set instance of k_seg for launches with few warps per CTA: parity, then cfg4 and the shard shapes with / without it
mkdir -p gpurun_out; rm -f gpurun_out/diag_l.jsonl
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -x -k "seg or cfg4 or format" > gpurun_out/pytest_l.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_l.log
for F in 1 0; do
  VIREO_B200_SEG_FEW=$F timeout 300 python scripts/time_passes.py cfg4 10 "cfg4 few=$F" 2>&1 | tail -1 | tee -a gpurun_out/diag_l.jsonl | cut -c1-200
  for N in 4 8; do VIREO_B200_SEG_FEW=$F timeout 300 python scripts/time_shard.py $N 10 2>&1 | tail -1 | tee -a gpurun_out/diag_l.jsonl | cut -c1-200; done
done
