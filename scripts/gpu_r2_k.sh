#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/diag_k.jsonl
for N in 2 4 8; do timeout 300 python scripts/time_shard.py $N 10 2>&1 | tail -1 | tee -a gpurun_out/diag_k.jsonl | cut -c1-400; done
for P in rows seg; do VIREO_B200_PATH=$P timeout 300 python scripts/time_passes.py cfg2 20 "cfg2 $P" 2>&1 | tail -1 | tee -a gpurun_out/diag_k.jsonl | cut -c1-300; done
