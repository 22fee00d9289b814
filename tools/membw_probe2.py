"""Write-only HBM bandwidth on B200 with constant and with incompressible data (is the 7.5 TB/s fill figure a
zero-fill artefact?).  torch built-ins only; used to interpret the roofline of the plane-store kernels."""
import torch
n = 1 << 32   # 16 GiB of float32
x = torch.empty(n, dtype=torch.float32, device="cuda")
row = torch.randn(1 << 20, dtype=torch.float32, device="cuda")          # 4 MB, L2 resident
xv = x.view(-1, 1 << 20)
def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(reps):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
t = timeit(lambda: x.zero_()); print(f"write-only zero            : {4*n/t/1e6:8.1f} GB/s")
t = timeit(lambda: x.fill_(1.2345)); print(f"write-only constant        : {4*n/t/1e6:8.1f} GB/s")
t = timeit(lambda: xv.copy_(row)); print(f"write-only random (bcast)  : {4*n/t/1e6:8.1f} GB/s")
t = timeit(lambda: torch.arange(n, out=x)); print(f"write-only arange          : {4*n/t/1e6:8.1f} GB/s")
y = torch.empty(n // 2, dtype=torch.float32, device="cuda")
t = timeit(lambda: y.copy_(x[: n // 2])); print(f"copy (r+w)                 : {4*n/t/1e6:8.1f} GB/s")
