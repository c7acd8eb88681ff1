"""Summarise an `ncu --page source --csv` export: per kernel, the instructions with most stall samples and
the stall-reason totals.  usage: ncu_src.py file.csv [kernel_index] [topN]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
# split into kernels
blocks = []; cur = None
for r in rows:
    if r and r[0] == "Kernel Name": cur = {"name": r[1], "rows": []}; blocks.append(cur); continue
    if cur is not None: cur["rows"].append(r)
b = blocks[which]
hdr = b["rows"][0]; data = b["rows"][1:]
ix = {h: i for i, h in enumerate(hdr)}
print(b["name"], len(data), "instructions")
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = collections.Counter()
for r in data:
    for h in stall_cols:
        try: tot[h] += int(r[ix[h]])
        except: pass
S = sum(tot.values())
print("stall totals:", ", ".join("%s %.1f%%" % (k[6:], 100.0 * v / S) for k, v in tot.most_common(10)))
samp = ix["# Samples"]; ex = ix["Instructions Executed"]
tot_inst = sum(int(r[ex]) for r in data)
print("total warp instructions", tot_inst, "samples", sum(int(r[samp]) for r in data))
order = sorted(range(len(data)), key=lambda i: -int(data[i][samp]))[:top]
for i in sorted(order):
    r = data[i]
    st = sorted(((int(r[ix[h]]), h[6:]) for h in stall_cols), reverse=True)[:3]
    print("%5d %-58s samp %6s exec %9s thr %5s  %s" % (i, r[ix["Source"]].strip()[:58], r[samp], r[ex], r[ix["Avg. Predicated-On Threads Executed"]],
          " ".join("%s:%d" % (n, v) for v, n in st if v)))
