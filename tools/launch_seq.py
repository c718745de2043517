#!/usr/bin/env python
"""Per-launch view of an ncu launch list (gpu__time_duration.sum, launch__grid_size, launch__block_size): prints the last
step's launches above a threshold and the per-kernel totals.   python tools/launch_seq.py gpurun_out/launches.csv [min_us] [nsteps]"""
import collections, csv, re, sys
src = sys.argv[1]
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 12.0
nsteps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
rows = list(csv.DictReader(l for l in open(src) if not l.startswith("==")))
byid = collections.OrderedDict()
for r in rows:
    d = byid.setdefault(r["ID"], {"name": re.sub(r"\(.*", "", r["Kernel Name"])})
    d[r["Metric Name"]] = (float(r["Metric Value"].replace(",", "")), r["Metric Unit"])
ids = list(byid)
ids = ids[len(ids) - len(ids) // nsteps:]
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
for k, i in enumerate(ids):
    d = byid[i]
    t, u = d["gpu__time_duration.sum"]
    t = t / 1e3 if u == "ns" else (t * 1e3 if u == "ms" else t)
    g = int(d.get("launch__grid_size", (0,))[0]); b = int(d.get("launch__block_size", (0,))[0])
    n = d["name"].replace("void ", "").replace("faln::<unnamed>::", "F:").replace("at::", "")
    agg[n[:80]][0] += 1; agg[n[:80]][1] += t; tot += t
    if t >= thr:
        print(f"{k:4d} {t:8.1f} g={g:6d} b={b:4d} {n[:90]}")
print(f"--- one step: {tot:.1f} us over {len(ids)} launches")
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
    print(f"{t:9.1f} {c:4d} {100 * t / tot:5.1f}%  {k}")
