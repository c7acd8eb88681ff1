#!/bin/bash
# which launches make racecheck report on the shipped build: the four-landing-set instance or the long-table workload?
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
for F in 0 1; do
  VIREO_B200_SEG_FEW=$F VIREO_B200_SAN_NOSPLIT=1 timeout 600 $CS --tool racecheck --racecheck-report analysis --print-limit 3 python scripts/sanitize_run.py seg small > gpurun_out/racecheck_few$F.log 2>&1
  echo "FEW=$F (no long-table case): $(grep 'RACECHECK SUMMARY' gpurun_out/racecheck_few$F.log)"
done
VIREO_B200_SEG_FEW=0 timeout 600 $CS --tool racecheck --racecheck-report analysis --print-limit 3 python scripts/sanitize_run.py seg small > gpurun_out/racecheck_few0_long.log 2>&1
echo "FEW=0 with the long-table case: $(grep 'RACECHECK SUMMARY' gpurun_out/racecheck_few0_long.log)"
VIREO_B200_SEG_FEW=1 VIREO_B200_SAN_NOSPLIT=1 timeout 600 $CS --tool racecheck --racecheck-report hazard --print-limit 4 python scripts/sanitize_run.py seg small > gpurun_out/racecheck_few1_hazard.log 2>&1
grep -m6 -A3 "hazard detected\|Warning: Race\|Error: Race\|WAR\|RAW" gpurun_out/racecheck_few1_hazard.log | head -40 | cut -c1-220
