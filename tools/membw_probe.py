"""Probe of B200 HBM bandwidth by direction (torch built-ins; used only to interpret rooflines)."""
import torch
n = 1 << 32   # 16 GiB of float32
x = torch.empty(n, dtype=torch.float32, device="cuda")
y = torch.empty(n, dtype=torch.float32, device="cuda")
def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(reps):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
t = timeit(lambda: x.fill_(1.0)); print(f"write-only fill : {4*n/t/1e6:8.1f} GB/s")
t = timeit(lambda: y.copy_(x));   print(f"copy (r+w)      : {8*n/t/1e6:8.1f} GB/s")
t = timeit(lambda: x.sum());      print(f"read-only sum   : {4*n/t/1e6:8.1f} GB/s")
t = timeit(lambda: torch.add(x, 1.0, out=y)); print(f"add (r+w)       : {8*n/t/1e6:8.1f} GB/s")
