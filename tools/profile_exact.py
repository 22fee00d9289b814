"""
One call of the exact (FFT) CWT + information path at a reduced channel count, for ncu captures of its kernels:

    python tools/profile_exact.py float64 2          # must exit 0 first
    ncu --set full --clock-control none --import-source on -k regex:'cwtf_|info_plane|fft_pass' --launch-skip N \
        --launch-count M -o gpurun_out/r02_exact_f64 python tools/profile_exact.py float64 2
"""
import os
import sys

import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from bench import FS, LOG2_N, ORDER, synth_batch_torch  # noqa: E402
from quantum_inferno_b200 import cwt_entropy  # noqa: E402

dtype = sys.argv[1] if len(sys.argv) > 1 else "float64"
chans = int(sys.argv[2]) if len(sys.argv) > 2 else 2
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
dev = torch.device("cuda", 0)
n = 1 << LOG2_N
x = synth_batch_torch(torch, n, list(range(chans)), dev, dtype)
nb = len(cwt_entropy.scales.log_frequency_hz_from_fft_points(FS, n, ORDER))
tdt = getattr(torch, dtype)
power = torch.empty(chans, nb, n, dtype=tdt, device=dev)
info = torch.empty_like(power)
for _ in range(reps):
    r = cwt_entropy.cwt_power_entropy(ORDER, x, FS, dtype=dtype, out_power=power, out_info=info, method="exact")
torch.cuda.synchronize()
print("entropy bits ch0:", float(r.entropy_bits()[0].item()))
