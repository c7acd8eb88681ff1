#!/bin/bash
# ncu --set full of the segment kernels of the shipped build (2 launches: SNP pass, cell pass); only summaries come back
mkdir -p gpurun_out /tmp/ncu
ncu --set full --clock-control none --import-source on -k regex:'k_seg' -s 4 -c 2 -f -o /tmp/ncu/prof_seg5 python scripts/time_passes.py cfg3 4 x > gpurun_out/ncu_seg5.log 2>&1; echo "ncu rc=$?"
ncu -i /tmp/ncu/prof_seg5.ncu-rep --page raw --csv > gpurun_out/prof_seg5_raw.csv 2>/dev/null
ncu -i /tmp/ncu/prof_seg5.ncu-rep --page source --csv > /tmp/ncu/prof_seg5_source.csv 2>/dev/null
python scripts/ncu_pick.py gpurun_out/prof_seg5_raw.csv > gpurun_out/ncu_seg5_pick.txt 2>&1
for K in 0 1; do python scripts/ncu_src.py /tmp/ncu/prof_seg5_source.csv $K 60 > gpurun_out/ncu_seg5_src$K.txt 2>&1; done
cp /tmp/ncu/prof_seg5.ncu-rep gpurun_out/
grep -E "==|time_duration|wavefronts_mem_shared|inst_executed.sum|issue_active|dram__bytes|sleeping|long_sc|short_sc" gpurun_out/ncu_seg5_pick.txt | cut -c1-140
