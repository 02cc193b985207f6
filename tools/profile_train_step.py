"""Kernel census of one AIR-ASR training step (eager) at a given local batch: python tools/profile_train_step.py 512"""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from mog_asr_b200.air import Trainer, config_from_flags, bench_train
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
dev = torch.device("cuda:0")
cfg = config_from_flags("mnist", "24", always_max_steps=True)
tr = Trainer(cfg, dev, global_batch=B)
x = bench_train.synthetic_batch(cfg, B, 0, dev)
for _ in range(3): tr.step(x)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    tr.step(x); torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
agg = collections.defaultdict(lambda: [0, 0.0])
for e in ev:
    a = agg[e.name[:70]]; a[0] += 1; a[1] += e.device_time if hasattr(e, "device_time") else e.cuda_time
tot = sum(v[1] for v in agg.values()); n = sum(v[0] for v in agg.values())
print(f"batch {B}: {n} kernels, {tot/1e3:.2f} ms of GPU time")
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:80]:
    print(f"{c:5d} {t/1e3:8.3f} ms {100*t/tot:5.1f}%  {k}")
ops = collections.Counter()
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CPU and e.name.startswith("aten::") and e.cpu_parent is not None and not e.cpu_parent.name.startswith("aten::"):
        ops[e.name] += 1
print("top-level aten ops:", sum(ops.values()))
for k, c in ops.most_common(40):
    print(f"{c:5d} {k}")
