"""Render the result tables of BASELINE.md section 3 from the JSON files under profiles/.
usage: python tools/make_tables.py r01"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
b = json.load(open(os.path.join(ROOT, "profiles", f"{tag}_bench_n1.json")))
sw = json.load(open(os.path.join(ROOT, "profiles", f"sweep_{tag}.json")))
out = []
k = b["roofline"]["kernels"]
cb = b.get("cpu_baseline", {})
out.append(f"### 3.1 Headline cell -- 1x B200\n\n{b['config']['workload']}\n")
out.append("| quantity | value |\n|---|---|")
out.append(f"| glimpses/s, inputs resident (value) | {b['value']/1e6:.1f} M ({b['ms_per_step']:.2f} ms per step of 262 144 glimpses) |")
if "e2e" in b:
    out.append(f"| glimpses/s end to end, host buffers (e2e) | {b['e2e']['value']/1e6:.2f} M ({b['e2e']['h2d_bytes_per_step']/1e9:.2f} GB H2D + {b['e2e']['d2h_bytes_per_step']/1e9:.2f} GB D2H per step) |")
if cb:
    out.append(f"| CPU port (oracle/stn_ref.c, {cb['cores']} threads) | {cb['value']/1e6:.3f} M glimpses/s |")
out.append(f"| step algorithmic GB/s / fraction of {b['roofline']['peak']:.1f} GB/s | {b['roofline']['step_alg_gbs']:.0f} / {b['roofline']['step_frac']:.2f} |")
out.append(f"| SM clock during the run | {b['clocks']['sm_mhz']:.0f} MHz (max {b['clocks']['sm_max_mhz']:.0f}), reasons {b['clocks']['reasons']} |\n")
out.append("| kernel | us per launch | algorithmic MB | GB/s | fraction of peak |\n|---|---|---|---|---|")
for n, v in k.items():
    out.append(f"| {n} | {v['ms']*1e3:.1f} | {v['alg_bytes_per_launch']/1e6:.1f} | {v['achieved_gbs']:.0f} | {v['frac']:.2f} |")
out.append(f"\n### 3.2 Config 5 sweep, 1x B200, B = 16 384, 8 AIR steps (read + write, fwd + bwd dU+dtheta); fraction of {sw['peak_gbs']:.1f} GB/s by algorithmic bytes\n")
out.append("| canvas | glimpse | theta | ms/step | M glimpses/s | step frac | read fwd us (frac) | read bwd us (frac) | write fwd us (frac) | write bwd us (frac) |\n|---|---|---|---|---|---|---|---|---|---|")
for r in sw["cells"]:
    kk = r["kernels"]
    f = lambda n: f"{kk[n]['us']:.0f} ({kk[n]['frac']:.2f})"
    out.append(f"| {r['canvas']} | {r['glimpse']} | {r['regime']} | {r['ms_per_step']:.2f} | {r['glimpses_per_sec']/1e6:.1f} | {r['step_frac']:.2f} | {f('read_fwd')} | {f('read_bwd')} | {f('write_fwd')} | {f('write_bwd')} |")
if "train" in b:
    out.append("\n### 3.3 AIR-ASR training step (configs 2-4), 1x B200\n")
    out.append("| run | global batch | ms/step | images/s | loop steps | mode |\n|---|---|---|---|---|---|")
    for n, v in b["train"].items():
        if "images_per_sec" in v:
            out.append(f"| {n} | {v['global_batch']} | {v['ms_per_step']:.1f} | {v['images_per_sec']:.0f} | {v.get('mean_loop_steps','')} | {v.get('mode', v.get('sample',''))} |")
print("\n".join(out))
