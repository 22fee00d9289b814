"""
Quick timing of the headline step (bench.py's north-star workload) with the per-category CUDA-event breakdown of the
library -- no CPU baseline, no end-to-end leg, no clock sampling: for A/B comparisons of kernel changes.

    python tools/quick_step.py [steps] [label]
    QI_L2K_SCALAR=1 python tools/quick_step.py 10 scalar

Prints one JSON line: ms per step, category ms per step, entropy of channel 0, and the relative L2 distance of three
power rows (a level-0 band, the top band, a deep band) from the exact float32 method (full-length FFT route).
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from bench import CH_PER_GPU, FS, LOG2_N, ORDER, synth_batch_torch  # noqa: E402
from quantum_inferno_b200 import _lib, cwt_entropy  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
label = sys.argv[2] if len(sys.argv) > 2 else ""
dev = torch.device("cuda", 0)
n = 1 << LOG2_N
x = synth_batch_torch(torch, n, list(range(CH_PER_GPU)), dev)
nb = len(cwt_entropy.scales.log_frequency_hz_from_fft_points(FS, n, ORDER))
power = torch.empty(CH_PER_GPU, nb, n, dtype=torch.float32, device=dev)
info = torch.empty_like(power)
from quantum_inferno_b200._runtime import get_runtime  # noqa: E402
lib = get_runtime(x).lib


def step():
    return cwt_entropy.cwt_power_entropy(ORDER, x, FS, dtype="float32", out_power=power, out_info=info,
                                         method="multirate")


for _ in range(3):
    r = step()
torch.cuda.synchronize()
lib.qi_profile_enable(1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    r = step()
e1.record()
torch.cuda.synchronize()
lib.qi_profile_enable(0)
cat_ms = (np.zeros(_lib.QI_N_CATEGORIES), np.zeros(_lib.QI_N_CATEGORIES, dtype=np.int64))
lib.qi_profile_read(cat_ms[0].ctypes.data, cat_ms[1].ctypes.data)
out = {"label": label, "ms_per_step": e0.elapsed_time(e1) / steps,
       "category_ms": {nm: round(float(ms) / steps, 4) for nm, ms in zip(_lib.CATEGORY_NAMES, cat_ms[0])},
       "entropy_bits_ch0": float(r.entropy_bits()[0].item())}
if os.environ.get("QI_QUICK_CHECK", "1") != "0":
    # one channel, exact float32 method on three rows
    rows = [nb - 1, nb - 5, nb - 7, nb - 8, nb - 10, nb - 11, 20]
    ex = cwt_entropy.cwt_power_entropy(ORDER, x[:1], FS, dtype="float32", method="exact", want_info=False)
    l2 = {}
    for b in rows:
        a, e = power[0, b].double(), ex.power[0, b].double()
        l2[str(b)] = float(((a - e).norm() / e.norm()).item())
    out["power_l2_vs_exact"] = l2
print(json.dumps(out))
