"""
Micro-benchmark of the multi-pass FFT (qi_fft_c2c): forward + inverse of [batch, 2^m] complex records resident in HBM,
CUDA-event time per direction, algorithmic traffic (one read + one write of the record per direction) in GB/s.

    python tools/fft_probe.py            # prints one JSON line per (dtype, batch, log2n)
"""
import json
import os
import sys

import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from quantum_inferno_b200 import _lib, _runtime  # noqa: E402

rt = _runtime.get_runtime()
lib = rt.lib
CASES = [("float32", 8, 25), ("float64", 8, 25), ("float32", 16, 18), ("float64", 16, 18), ("float32", 672, 18),
         ("float64", 24, 20), ("float32", 4096, 10), ("float64", 4096, 10), ("float32", 1024, 14), ("float64", 1024, 14)]
if len(sys.argv) > 1:
    CASES = [(sys.argv[1], int(sys.argv[2]), int(sys.argv[3]))]
for dt, batch, m in CASES:
    cdt = torch.complex64 if dt == "float32" else torch.complex128
    n = 1 << m
    x = torch.randn(batch, n, dtype=cdt, device=rt.device)
    y = torch.empty_like(x)
    z = torch.empty_like(x)
    code = _runtime.DTYPE_CODE[dt]
    out = {"dtype": dt, "batch": batch, "log2n": m}
    for name, src, dst, inv in (("fwd", x, y, 0), ("inv", y, z, 1)):
        for _ in range(2):
            assert lib.qi_fft_c2c(rt.ptr(src), rt.ptr(dst), batch, m, inv, code, rt.stream()) == 0
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        e0.record()
        for _ in range(reps):
            lib.qi_fft_c2c(rt.ptr(src), rt.ptr(dst), batch, m, inv, code, rt.stream())
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        out[name + "_ms"] = ms
        out[name + "_alg_GBps"] = 2 * x.numel() * x.element_size() / ms / 1e6
    err = float((z - x).abs().max() / x.abs().max())        # the inverse is scaled by 1 / n
    out["roundtrip_max_rel_err"] = err
    print(json.dumps(out), flush=True)
    del x, y, z
    torch.cuda.empty_cache()
