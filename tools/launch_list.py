#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (launches, total, mean, share of sampler time).

    python tools/launch_list.py gpurun_out/launches_r02_256x64.csv "header comment" > profiles/r02_stn_c5_256x64_launch_list.txt
"""
import collections
import csv
import sys

path, note = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
rows = list(csv.reader(open(path)))
hdr = next(r for r in rows if r and r[0] == "ID")
ix = {h: i for i, h in enumerate(hdr)}
scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6}
agg = collections.OrderedDict()
for r in rows:
    if not r or not r[0].isdigit() or r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    us = float(r[ix["Metric Value"]].replace(",", "")) * scale[r[ix["Metric Unit"]]]
    a = agg.setdefault(r[ix["Kernel Name"]], [0, 0.0])
    a[0] += 1
    a[1] += us
sampler = sum(v[1] for k, v in agg.items() if "stn_fwd" in k or "stn_bwd" in k)
print(f"# {note}")
print("# First launches of the process; cold-cache, serialised -- compare SHARES.  Sampler kernels only in the share column")
print("# (torch fill / rng kernels build the inputs).")
print(f"# {'kernel':<70} {'launches':>8} {'total us':>12} {'mean us':>10} share of sampler time")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    share = f"{100 * t / sampler:5.1f}%" if ("stn_fwd" in k or "stn_bwd" in k) and sampler else "    -"
    print(f"{k[:72]:<72} {n:>8} {t:>12.1f} {t / n:>10.2f}   {share}")
