#!/bin/bash
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
timeout 900 $CS --tool memcheck --print-limit 20 python scripts/sanitize_run.py seg small > gpurun_out/memcheck_seg_final.log 2>&1; echo "memcheck rc=$?"; grep -E "heavy rows|split launches|ERROR SUMMARY" gpurun_out/memcheck_seg_final.log
timeout 900 $CS --tool initcheck --print-limit 20 python scripts/sanitize_run.py seg small > gpurun_out/initcheck_seg_final.log 2>&1; echo "initcheck rc=$?"; grep -E "ERROR SUMMARY" gpurun_out/initcheck_seg_final.log
timeout 900 $CS --tool synccheck --print-limit 20 python scripts/sanitize_run.py seg small > gpurun_out/synccheck_seg_final.log 2>&1; echo "synccheck rc=$?"; grep -E "ERROR SUMMARY" gpurun_out/synccheck_seg_final.log
