#!/bin/bash
# run on the GPU box via gpurun: bench + ncu launch list + one full ncu capture of the gather kernels
mkdir -p gpurun_out
W=${1:-cfg3}
python bench.py --workload $W --steps 3 --warmup 3 > gpurun_out/bench_$W.json 2> gpurun_out/bench_$W.err; echo "bench rc=$?"
cat gpurun_out/bench_$W.json; tail -5 gpurun_out/bench_$W.err
if [ "${2:-}" = "ncu" ]; then
  ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 120 --csv --log-file gpurun_out/launches_$W.csv python bench.py --workload $W --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_bench.log 2>&1; echo "ncu list rc=$?"
  ncu --set full --clock-control none --import-source on -k regex:'k_gather' -s 4 -c 2 -f -o gpurun_out/prof_$W python bench.py --workload $W --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
  ncu -i gpurun_out/prof_$W.ncu-rep --page raw --csv > gpurun_out/prof_${W}_raw.csv 2>/dev/null
  ncu -i gpurun_out/prof_$W.ncu-rep --page details --csv > gpurun_out/prof_${W}_details.csv 2>/dev/null
  ncu -i gpurun_out/prof_$W.ncu-rep --page source --csv > gpurun_out/prof_${W}_source.csv 2>/dev/null
  ls -la gpurun_out
fi
