#!/bin/bash
mkdir -p gpurun_out
for V in _sleeppc _sleepa _sleepb _sleepc; do
  VIREO_B200_LIB=$PWD/vireo_b200/libvireo_b200${V}.so timeout 300 python scripts/time_passes.py cfg3 20 "variant${V}" 2>&1 | tail -1 | tee -a gpurun_out/diag_q.jsonl | cut -c1-200
done
VIREO_B200_LIB=$PWD/vireo_b200/libvireo_b200_sleeppc.so timeout 300 python scripts/time_passes.py cfg4 20 "cfg4 sleeppc" 2>&1 | tail -1 | tee -a gpurun_out/diag_q.jsonl | cut -c1-200
timeout 300 python scripts/time_passes.py cfg4 20 "cfg4 default" 2>&1 | tail -1 | tee -a gpurun_out/diag_q.jsonl | cut -c1-200
