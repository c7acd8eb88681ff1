#!/bin/bash
# round 2: record chunks staged in shared memory -- format check, parity of the segment families, timings, ring geometry
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fullsize.py -m gpu -q -x -k "segment_format" > gpurun_out/pytest_g0.log 2>&1
rc=$?; echo "format rc=$rc"; tail -3 gpurun_out/pytest_g0.log
if [ $rc -ne 0 ]; then grep -n "Error\|assert" gpurun_out/pytest_g0.log | head -20; exit 1; fi
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "seg" > gpurun_out/pytest_g1.log 2>&1
rc=$?; echo "parity rc=$rc"; tail -3 gpurun_out/pytest_g1.log
if [ $rc -ne 0 ]; then grep -n "Error\|assert" gpurun_out/pytest_g1.log | head -20; exit 1; fi
L=vireo_b200/libvireo_b200
rm -f gpurun_out/diag_g.jsonl
for V in "" _nolds; do
  VIREO_B200_LIB=$PWD/${L}${V}.so timeout 300 python scripts/time_passes.py cfg3 10 "variant${V}" 2>&1 | tail -1 | tee -a gpurun_out/diag_g.jsonl | cut -c1-300
done
for G in "384 4 3" "192 8 6" "512 3 2" "128 12 9" "256 6 5"; do
  set -- $G
  VIREO_B200_SEG_WR64=$1 VIREO_B200_SEG_NB64=$2 VIREO_B200_SEG_SPAN64=$3 timeout 300 python scripts/time_passes.py cfg3 10 "win=$1 nb=$2 span=$3" 2>&1 | tail -1 | tee -a gpurun_out/diag_g.jsonl | cut -c1-300
done
timeout 300 python scripts/time_passes.py cfg4 10 "cfg4" 2>&1 | tail -1 | tee -a gpurun_out/diag_g.jsonl | cut -c1-300
VIREO_B200_SEG_WR32=1024 VIREO_B200_SEG_NB32=3 VIREO_B200_SEG_SPAN32=2 timeout 300 python scripts/time_passes.py cfg4 10 "cfg4 1024 3 2" 2>&1 | tail -1 | tee -a gpurun_out/diag_g.jsonl | cut -c1-300
