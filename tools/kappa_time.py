"""Times the fused multirate path at the level-assignment half-widths 4.8 / 2.4 (`_plan.MR_KAPPA`) for three configurations
(order 12 half table of a 2^25 record, order 6 8 x 2^22, order 3 8 x 2^24); measurement tool."""
import sys, torch, numpy as np
import os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from bench import FS, synth_batch_torch
from quantum_inferno_b200 import cwt_entropy, _plan
dev = torch.device("cuda", 0)
for order, logn, ch in ((12, 25, 1), (6, 22, 8), (3, 24, 8)):
    n = 1 << logn
    x = synth_batch_torch(torch, n, list(range(ch)), dev)
    nb = len(cwt_entropy.scales.log_frequency_hz_from_fft_points(FS, n, order))
    if order == 12:
        bs = (0, nb // 2)
    else:
        bs = None
    nbl = nb if bs is None else bs[1] - bs[0]
    power = torch.empty(ch, nbl, n, dtype=torch.float32, device=dev); info = torch.empty_like(power)
    for ka in (4.8, 2.4):
        _plan.MR_KAPPA = ka
        f = cwt_entropy.scales.log_frequency_hz_from_fft_points(FS, n, order)
        lv = _plan.multirate_bands(order, n, f, FS)[0]['level']
        for _ in range(3):
            r = cwt_entropy.cwt_power_entropy(order, x, FS, dtype="float32", out_power=power, out_info=info, method="multirate", band_slice=bs)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            r = cwt_entropy.cwt_power_entropy(order, x, FS, dtype="float32", out_power=power, out_info=info, method="multirate", band_slice=bs)
        e1.record(); torch.cuda.synchronize()
        print(order, logn, ch, 'kappa', ka, 'ms', e0.elapsed_time(e1) / 5, 'levels', np.bincount(lv).tolist(), flush=True)
