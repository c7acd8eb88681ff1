#!/bin/bash
# ncu --set full capture of the window-segment kernels (cell pass + SNP pass) for one family
mkdir -p gpurun_out
P=${1:-seg}
W=${2:-cfg3}
VIREO_B200_PATH=$P ncu --set full --clock-control none --import-source on -k regex:'k_seg' -s 4 -c 2 -f -o gpurun_out/prof_$P python bench.py --workload $W --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_$P.log 2>&1; echo "ncu full rc=$?"
ncu -i gpurun_out/prof_$P.ncu-rep --page raw --csv > gpurun_out/prof_${P}_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_$P.ncu-rep --page source --csv > gpurun_out/prof_${P}_source.csv 2>/dev/null
ls -la gpurun_out | tail -8
