#!/bin/bash
# window waits with the warp asleep between polls (producer / producer + consumers) against plain try_wait spinning
mkdir -p gpurun_out; rm -f gpurun_out/diag_i.jsonl
for V in "" _sleepp _sleeppc; do
  VIREO_B200_LIB=$PWD/vireo_b200/libvireo_b200${V}.so timeout 300 python scripts/time_passes.py cfg3 10 "variant${V}" 2>&1 | tail -1 | tee -a gpurun_out/diag_i.jsonl | cut -c1-230
done
VIREO_B200_LIB=$PWD/vireo_b200/libvireo_b200_sleeppc.so timeout 300 python scripts/time_passes.py cfg4 10 "cfg4 sleeppc" 2>&1 | tail -1 | tee -a gpurun_out/diag_i.jsonl | cut -c1-230
timeout 300 python scripts/time_passes.py cfg4 10 "cfg4" 2>&1 | tail -1 | tee -a gpurun_out/diag_i.jsonl | cut -c1-230
