#!/usr/bin/env python
"""Per-kernel DRAM bytes and time of the last `n` launches of an ncu CSV (one wide-world tick).
Usage: python tools/wide_traffic_sum.py launches.csv 43"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2])
hdr = None
launches = collections.OrderedDict()
for r in rows:
    if "Kernel Name" in r:
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        v = float(d["Metric Value"].replace(",", ""))
        u = d["Metric Unit"]
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3}.get(u, 1.0)
        e = launches.setdefault(d["ID"], {"name": d["Kernel Name"].split("(")[0], "bytes": 0.0, "us": 0.0})
        if d["Metric Name"].startswith("dram__bytes"):
            e["bytes"] += v * scale
        elif d["Metric Name"] == "gpu__time_duration.sum":
            e["us"] += v * scale
last = list(launches.values())[-n:]
per = collections.OrderedDict()
for e in last:
    p = per.setdefault(e["name"], [0, 0.0, 0.0])
    p[0] += 1
    p[1] += e["bytes"]
    p[2] += e["us"]
tb, tu = sum(e["bytes"] for e in last), sum(e["us"] for e in last)
print(f"last {len(last)} launches (one tick): {tb / 1e6:.1f} MB DRAM read + written, {tu:.1f} us of kernel time (cold caches, serialised)")
for k, (c, b, u) in sorted(per.items(), key=lambda kv: -kv[1][2]):
    print(f"{k:28s} {c:3d} launches {b / 1e6:9.2f} MB {u:8.1f} us {100 * u / tu:5.1f} %")
