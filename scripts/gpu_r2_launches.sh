#!/bin/bash
# round 2: ncu launch list (per-launch device time, cold cache, serialised) of the headline bench command
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_cfg3.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-fixed32 --no-wrap > gpurun_out/ncu_bench_r02.log 2>&1; echo "ncu list rc=$?"
python - > gpurun_out/r02_launches_cfg3_summary.txt <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/r02_launches_cfg3.csv")) if len(r) > 5]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict(); tot = 0.0
for r in rows[1:]:
    try: v = float(r[vi].replace(",", ""))
    except ValueError: continue
    if r[ui] == "ns": v /= 1000.0
    elif r[ui] == "ms": v *= 1000.0
    a = agg.setdefault(r[ki][:64], [0, 0.0]); a[0] += 1; a[1] += v; tot += v
print("# ncu launch list (gpu__time_duration.sum, --clock-control none): python bench.py --steps 2 --warmup 1 --no-cpu --no-fixed32 --no-wrap")
print("# first 400 launches of the run: staging, format builds, warm-up and timed fits (cold-cache, serialised: shares, not absolutes,")
print("# compare with bench.py's CUDA-event times)")
print("%-66s %5s %12s %8s" % ("kernel", "n", "avg us", "share"))
for k, (n, t) in agg.items(): print("%-66s %5d %12.1f %7.1f%%" % (k, n, t / n, 100.0 * t / tot))
PY
cat gpurun_out/r02_launches_cfg3_summary.txt
