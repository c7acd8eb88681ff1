#!/bin/bash
# ncu --set full of the segment kernels: shipped build and the no-table-loads diagnostic build; only summaries come back
mkdir -p gpurun_out /tmp/ncu
for V in "" _nolds; do
  VIREO_B200_LIB=$PWD/vireo_b200/libvireo_b200${V}.so ncu --set full --clock-control none --import-source on -k regex:'k_seg' -s 4 -c 2 -f -o /tmp/ncu/prof_seg4${V} python scripts/time_passes.py cfg3 4 x > gpurun_out/ncu_seg4${V}.log 2>&1; echo "ncu rc=$?"
  ncu -i /tmp/ncu/prof_seg4${V}.ncu-rep --page raw --csv > gpurun_out/prof_seg4${V}_raw.csv 2>/dev/null
  ncu -i /tmp/ncu/prof_seg4${V}.ncu-rep --page source --csv > /tmp/ncu/prof_seg4${V}_source.csv 2>/dev/null
  python scripts/ncu_pick.py gpurun_out/prof_seg4${V}_raw.csv > gpurun_out/ncu_seg4${V}_pick.txt 2>&1
  for K in 0 1; do python scripts/ncu_src.py /tmp/ncu/prof_seg4${V}_source.csv $K 70 > gpurun_out/ncu_seg4${V}_src$K.txt 2>&1; done
done
cp /tmp/ncu/prof_seg4.ncu-rep gpurun_out/
ls -la gpurun_out | grep seg4
