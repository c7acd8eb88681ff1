#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/diag_y.jsonl
for S in 1 0; do VIREO_B200_SEG_OWNER_SPLIT=$S timeout 600 python scripts/time_heavy.py 16 2>&1 | tail -1 | tee -a gpurun_out/diag_y.jsonl | cut -c1-400; done
