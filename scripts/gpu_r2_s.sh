#!/bin/bash
# ring geometry once more, with the sleeping waits
mkdir -p gpurun_out; rm -f gpurun_out/diag_s.jsonl
for G in "256 6 4" "256 6 5" "192 8 6" "192 8 7" "384 4 3" "128 12 10"; do
  set -- $G
  VIREO_B200_SEG_WR64=$1 VIREO_B200_SEG_NB64=$2 VIREO_B200_SEG_SPAN64=$3 timeout 300 python scripts/time_passes.py cfg3 20 "win=$1 nb=$2 span=$3" 2>&1 | tail -1 | tee -a gpurun_out/diag_s.jsonl | cut -c1-230
done
