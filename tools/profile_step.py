"""
Runs `reps` steps of the headline workload (bench.py's north-star configuration, no timing, no CPU baseline) so that
ncu can be pointed at single kernels of one step:

    python tools/profile_step.py                                   # must exit 0 first
    ncu --set full --clock-control none --import-source on -k regex:'mr_level4_kernel|mr_expand_kernel' \
        --launch-skip 20 --launch-count 20 -o gpurun_out/r01_mr_full python tools/profile_step.py
"""
import os
import sys

import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from bench import CH_PER_GPU, FS, LOG2_N, METHOD, ORDER, synth_batch_torch  # noqa: E402
from quantum_inferno_b200 import cwt_entropy  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
dev = torch.device("cuda", 0)
n = 1 << LOG2_N
x = synth_batch_torch(torch, n, list(range(CH_PER_GPU)), dev)
nb = len(cwt_entropy.scales.log_frequency_hz_from_fft_points(FS, n, ORDER))
power = torch.empty(CH_PER_GPU, nb, n, dtype=torch.float32, device=dev)
info = torch.empty_like(power)
for _ in range(reps):
    r = cwt_entropy.cwt_power_entropy(ORDER, x, FS, dtype="float32", out_power=power, out_info=info, method=METHOD)
torch.cuda.synchronize()
print("entropy bits ch0:", float(r.entropy_bits()[0].item()))
