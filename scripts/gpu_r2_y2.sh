#!/bin/bash
mkdir -p gpurun_out
for G in "1.5 1.3" "1.5 1.0" "1.2 1.1"; do set -- $G
  VIREO_B200_SEG_SKEW=$1 VIREO_B200_SEG_CAP=$2 timeout 300 python scripts/time_passes.py cfg3 20 "skew=$1 cap=$2" 2>&1 | tail -1 | cut -c1-330
done
