"""Render the result tables of BASELINE.md section 3 from the JSON files under profiles/.
usage: python tools/make_tables.py r02"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
P = lambda n: os.path.join(ROOT, "profiles", n)
b = json.load(open(P(f"{tag}_bench_n1.json")))
sw = json.load(open(P(f"sweep_{tag}.json")))
out = []
k = b["roofline"]["kernels"]
cb = b.get("cpu_baseline", {})
rf = b["roofline"]
out.append(f"### 3.1 Headline cell -- 1x B200\n\n{b['config']['workload']}\n")
out.append("| quantity | value |\n|---|---|")
out.append(f"| glimpses/s, inputs resident (value) | {b['value']/1e6:.2f} M ({b['ms_per_step']:.2f} ms per step of 262 144 glimpses) |")
if "e2e" in b:
    out.append(f"| glimpses/s end to end, host buffers (e2e) | {b['e2e']['value']/1e6:.3f} M ({b['e2e']['h2d_bytes_per_step']/1e9:.2f} GB H2D + {b['e2e']['d2h_bytes_per_step']/1e9:.2f} GB D2H per step, {b['e2e']['ms_per_step']:.0f} ms) |")
if cb:
    out.append(f"| CPU port (oracle/stn_ref.c, {cb['cores']} threads) | {cb['value']/1e6:.4f} M glimpses/s; 1 thread: {cb.get('value_1thread', float('nan'))/1e6:.5f} M |")
out.append(f"| step, algorithmic GB/s / fraction of {rf['peak']:.1f} GB/s | {rf['step_alg_gbs']:.0f} / {rf['step_frac']:.2f} |")
if "step_frac_physical" in rf:
    out.append(f"| step, measured DRAM GB/s / fraction | {rf['step_physical_gbs']:.0f} / {rf['step_frac_physical']:.2f} |")
out.append(f"| SM clock during the run | {b['clocks']['sm_mhz']:.0f} MHz (max {b['clocks']['sm_max_mhz']:.0f}), reasons {b['clocks']['reasons']} |\n")
out.append("| kernel | us per launch | algorithmic MB | fraction (algorithmic) | DRAM MB (ncu) | fraction (physical) |\n|---|---|---|---|---|---|")
for n, v in k.items():
    tr = f"{v['traffic_bytes_per_launch']/1e6:.1f}" if "traffic_bytes_per_launch" in v else "-"
    fp = f"{v['frac_physical']:.2f}" if "frac_physical" in v else "-"
    out.append(f"| {n} | {v['ms']*1e3:.1f} | {v['alg_bytes_per_launch']/1e6:.1f} | {v['frac']:.2f} | {tr} | {fp} |")
out.append(f"\n### 3.2 Config 5 sweep, 1x B200, B = 16 384, 8 AIR steps (read + write, fwd + bwd dU+dtheta); fractions of {sw['peak_gbs']:.1f} GB/s: algorithmic / physical (ncu DRAM bytes)\n")
out.append("| canvas | glimpse | theta | ms/step | M glimpses/s | step frac alg / phys | read fwd us (a/p) | read bwd us (a/p) | write fwd us (a/p) | write bwd us (a/p) |\n|---|---|---|---|---|---|---|---|---|---|")
na = npy = 0
for r in sw["cells"]:
    kk = r["kernels"]
    f = lambda n: f"{kk[n]['us']:.0f} ({kk[n]['frac']:.2f}/{kk[n].get('frac_physical', float('nan')):.2f})"
    out.append(f"| {r['canvas']} | {r['glimpse']} | {r['regime']} | {r['ms_per_step']:.2f} | {r['glimpses_per_sec']/1e6:.1f} | {r['step_frac']:.2f} / {r.get('step_frac_physical', float('nan')):.2f} | {f('read_fwd')} | {f('read_bwd')} | {f('write_fwd')} | {f('write_bwd')} |")
    na += r["step_frac"] >= 0.6
    npy += r.get("step_frac_physical", 0) >= 0.6
out.append(f"\nCells at or above 0.60 of the measured peak: {npy} of {len(sw['cells'])} by physical bytes, {na} by algorithmic bytes.")
if "train" in b:
    out.append("\n### 3.3 AIR-ASR training step (configs 2-4), 1x B200\n")
    out.append("| run | global batch | ms/step | images/s | kernels per step | loop steps | mode |\n|---|---|---|---|---|---|---|")
    for n, v in b["train"].items():
        if isinstance(v, dict) and "images_per_sec" in v:
            out.append(f"| {n} | {v['global_batch']} | {v['ms_per_step']:.2f} | {v['images_per_sec']:.0f} | {v.get('kernels_per_step', '')} | {v.get('mean_loop_steps','')} | {v.get('mode', v.get('sample',''))} |")
rows = []
for n in (1, 2, 4, 8):
    try:
        rows.append(json.load(open(P(f"{tag}_bench_n{n}.json"))))
    except Exception:
        pass
if len(rows) > 1:
    v1 = rows[0]["value"]
    t1 = rows[0]["train"]["C4_global4096"]
    out.append(f"\n### 3.4 Multi-GPU (one process per GPU, batch shards; `profiles/{tag}_bench_n{{1,2,4,8}}.json`)\n")
    out.append("| GPUs | STN glimpses/s (headline cell, 16 384 canvases per GPU: weak scaling) | x vs 1 GPU | e2e glimpses/s | AIR-ASR C4 images/s, global batch 4096 (strong scaling) | AIR-ASR C4 images/s, 4096 per GPU (weak scaling) | dp_check max rel diff (all-reduced vs single-process gradient) |\n|---|---|---|---|---|---|---|")
    for r in rows:
        t = r["train"]
        s = t["C4_global4096"]
        w = t.get("C4_weak_4096_per_gpu", s)
        dc = t.get("dp_check", {})
        out.append(f"| {r['n_gpus']} | {r['value']/1e6:.2f} M | {r['value']/v1:.2f} | {r['e2e']['value']/1e6:.3f} M | {s['images_per_sec']/1e3:.0f} k ({s['ms_per_step']:.2f} ms/step; {s['images_per_sec']/t1['images_per_sec']:.2f}x) | {w['images_per_sec']/1e3:.0f} k ({w['ms_per_step']:.2f} ms/step; {w['images_per_sec']/t1['images_per_sec']:.2f}x) | {('%.1e' % dc['max_rel_diff']) if dc and 'max_rel_diff' in dc else '-'} |")
print("\n".join(out))
