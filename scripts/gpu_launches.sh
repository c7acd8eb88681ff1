#!/bin/bash
# ncu launch list (per-launch device time, cold cache) of one short bench run
mkdir -p gpurun_out
W=${1:-cfg3}
ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 80 --csv --log-file gpurun_out/launches_$W.csv python bench.py --workload $W --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_bench.log 2>&1; echo "ncu list rc=$?"
python - <<PY
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/launches_$W.csv")) if len(r) > 5]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    try: v = float(r[vi].replace(",", ""))
    except ValueError: continue
    a = agg.setdefault(r[ki][:70], [0, 0.0]); a[0] += 1; a[1] += v
for k, (n, t) in agg.items(): print("%-72s n=%3d  avg %.1f us" % (k, n, t / n / 1000.0 if "ns" in rows[1][hdr.index("Metric Unit")] else t / n))
PY
python - <<PY
import sys; sys.path.insert(0, ".")
import numpy as np, bench, vireo_b200 as vb
from vireo_b200 import _lib
AD, DP, w = bench.load_workload("$W")
c = vb.stage(AD, DP)
lib = _lib.load()
import ctypes as C
ws = _lib.WsSizes(); _lib.check(lib.vb_vireo_ws_sizes(c.handle, w["K"], 3, 1, 0, C.byref(ws)))
names = {9: "built", 10: "recA", 11: "recB", 12: "heavyA", 13: "bytes", 14: "gridA", 15: "gridB", 16: "lightA"}
print({v: int(lib.vb_counts_info(c.handle, k)) for k, v in names.items()})
PY
