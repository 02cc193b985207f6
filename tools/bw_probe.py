"""Measure write-only / read-only / copy HBM bandwidth with torch (context for the roofline fractions)."""
import torch, json
dev = torch.device("cuda:0")
n = 1 << 30  # 4 GiB of fp32
a = torch.empty(n, dtype=torch.float32, device=dev)
b = torch.empty(n, dtype=torch.float32, device=dev)
def timeit(f, reps=10):
    for _ in range(3): f()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
t = timeit(lambda: a.zero_());            w = 4 * n / t / 1e6
t = timeit(lambda: a.sum());              r = 4 * n / t / 1e6
t = timeit(lambda: b.copy_(a));           c = 8 * n / t / 1e6
t = timeit(lambda: a.add_(1.0));          rw = 8 * n / t / 1e6
print(json.dumps(dict(write_only_gbs=w, read_only_gbs=r, copy_gbs=c, inplace_rw_gbs=rw)))
