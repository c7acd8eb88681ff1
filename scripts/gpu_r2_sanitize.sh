#!/bin/bash
# round 2: compute-sanitizer evidence and the protocol canary.
#   default build   : table windows filled by cp.async.bulk (product)
#   canary build    : same, every released window buffer is overwritten with NaNs before its refill
#   plainfill build : canary + the window filled by plain loads / stores of the producer warp behind the same mbarriers
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
for V in default canary plainfill; do
  L=$PWD/vireo_b200/libvireo_b200_$V.so; [ $V = default ] && L=$PWD/vireo_b200/libvireo_b200.so
  VIREO_B200_LIB=$L python scripts/sanitize_run.py all > gpurun_out/san_digest_$V.txt 2>&1; echo "digest $V rc=$?"
done
cat gpurun_out/san_digest_*.txt
VIREO_B200_LIB=$PWD/vireo_b200/libvireo_b200_canary.so timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "seg" > gpurun_out/pytest_canary.log 2>&1; echo "pytest canary rc=$?"; tail -3 gpurun_out/pytest_canary.log
VIREO_B200_LIB=$PWD/vireo_b200/libvireo_b200_plainfill.so timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "seg" > gpurun_out/pytest_plainfill.log 2>&1; echo "pytest plainfill rc=$?"; tail -3 gpurun_out/pytest_plainfill.log
timeout 900 $CS --tool memcheck --print-limit 20 python scripts/sanitize_run.py all small > gpurun_out/memcheck_all.log 2>&1; echo "memcheck rc=$?"
tail -3 gpurun_out/memcheck_all.log
timeout 600 $CS --tool synccheck --print-limit 20 python scripts/sanitize_run.py all small > gpurun_out/synccheck_all.log 2>&1; echo "synccheck rc=$?"
tail -3 gpurun_out/synccheck_all.log
timeout 900 $CS --tool initcheck --print-limit 20 python scripts/sanitize_run.py all small > gpurun_out/initcheck_all.log 2>&1; echo "initcheck rc=$?"
tail -3 gpurun_out/initcheck_all.log
timeout 900 $CS --tool racecheck --racecheck-report analysis --print-limit 5 python scripts/sanitize_run.py seg small > gpurun_out/racecheck_default.log 2>&1; echo "racecheck default rc=$?"
tail -3 gpurun_out/racecheck_default.log
VIREO_B200_LIB=$PWD/vireo_b200/libvireo_b200_plainfill.so timeout 900 $CS --tool racecheck --racecheck-report analysis --print-limit 5 python scripts/sanitize_run.py seg small > gpurun_out/racecheck_plainfill.log 2>&1; echo "racecheck plainfill rc=$?"
tail -3 gpurun_out/racecheck_plainfill.log
timeout 900 $CS --tool racecheck --racecheck-report analysis --print-limit 5 python scripts/sanitize_run.py rows small > gpurun_out/racecheck_rows.log 2>&1; echo "racecheck rows rc=$?"
tail -3 gpurun_out/racecheck_rows.log
