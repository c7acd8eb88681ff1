#!/bin/bash
# round 2: where does the segment kernel's time go -- diagnostic builds (no table loads / no window hand-over), then ncu
mkdir -p gpurun_out
L=vireo_b200/libvireo_b200
for V in "" _nolds; do
  VIREO_B200_LIB=$PWD/${L}${V}.so timeout 300 python scripts/time_passes.py cfg3 10 "variant${V}" 2>/dev/null | tail -1 | tee -a gpurun_out/diag_f.jsonl | cut -c1-330
done
bash scripts/gpu_ncu_seg.sh seg cfg3
python scripts/ncu_pick.py gpurun_out/prof_seg_raw.csv > gpurun_out/ncu_seg_pick.txt 2>&1; head -50 gpurun_out/ncu_seg_pick.txt
