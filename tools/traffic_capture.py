#!/usr/bin/env python
"""Physical DRAM bytes per launch of the four sampler kernels on one sweep cell, for bench.py's `frac_physical`.

Two modes:
    python tools/traffic_capture.py run --canvas 256 --glimpse 64 --regime prior --batch 16384
        launches read_fwd, read_bwd, write_fwd, write_bwd (in this order) for two theta sets -- nothing else that
        matches `stn_(fwd|bwd)`.  Run it under
        ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \\
            -k regex:'stn_(fwd|bwd)' --csv --log-file gpurun_out/traffic.csv python tools/traffic_capture.py run ...
    python tools/traffic_capture.py parse gpurun_out/traffic.csv --canvas ... > profiles/traffic_latest.json
        averages the two launches of every kind and records the commit the capture was taken at.
    python tools/traffic_capture.py parse gpurun_out/traffic.csv --canvas ... --into profiles/traffic_r02_sweep.json
        adds the cell to a multi-cell file (keys "canvas:glimpse:regime") that `bench.py --sweep` reads.
"""
import argparse
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
KINDS = ("read_fwd", "read_bwd", "write_fwd", "write_bwd")


def main():
    p = argparse.ArgumentParser()
    p.add_argument("mode", choices=["run", "parse"])
    p.add_argument("csv", nargs="?")
    p.add_argument("--canvas", type=int, default=256)
    p.add_argument("--glimpse", type=int, default=64)
    p.add_argument("--regime", default="prior")
    p.add_argument("--batch", type=int, default=16384)
    p.add_argument("--into", default="")
    a = p.parse_args()
    if a.mode == "run":
        import torch
        import bench
        dev = torch.device("cuda", 0)
        torch.cuda.set_device(dev)
        bench.AIR_STEPS = 2
        wl = bench.GpuWorkload(a, dev, seed=10)
        for t in range(2):
            for k in KINDS:
                wl.launch(k, t)
        torch.cuda.synchronize(dev)
        return
    rows = [r for r in csv.reader(open(a.csv)) if r and r[0].isdigit()]
    hdr = next(r for r in csv.reader(open(a.csv)) if r and r[0] == "ID")
    ix = {h: i for i, h in enumerate(hdr)}
    per = {}
    for r in rows:
        per.setdefault(int(r[ix["ID"]]), {})[r[ix["Metric Name"]]] = (float(r[ix["Metric Value"]].replace(",", "")), r[ix["Metric Unit"]])
    ids = sorted(per)
    assert len(ids) == 8, f"expected 8 kernel launches, found {len(ids)}"
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    out = {k: 0.0 for k in KINDS}
    for n, i in enumerate(ids):
        m = per[i]
        b = sum(m[k][0] * scale[m[k][1]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        out[KINDS[n % 4]] += b / 2
    head = subprocess.run(["git", "rev-parse", "--short", "HEAD"], cwd=ROOT, capture_output=True, text=True).stdout.strip()
    rec = dict(source=f"ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none over tools/traffic_capture.py "
                      f"(two launches per kernel, averaged), taken at commit {head}",
               cell=dict(canvas=a.canvas, glimpse=a.glimpse, regime=a.regime, batch=a.batch),
               bytes_per_launch={k: int(v) for k, v in out.items()})
    if a.into:
        try:
            allc = json.load(open(a.into))
        except Exception:
            allc = {}
        allc[f"{a.canvas}:{a.glimpse}:{a.regime}"] = rec
        json.dump(allc, open(a.into, "w"), indent=1)
    else:
        json.dump(rec, sys.stdout, indent=1)


if __name__ == "__main__":
    main()
