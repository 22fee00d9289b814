"""Small run of every kernel family for `compute-sanitizer --tool memcheck python tools/sanitizer_smoke.py`
(out-of-bounds / misaligned accesses); sizes chosen so that partial tiles, odd band counts and halos are hit."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from quantum_inferno_b200 import cwt_atoms, cwt_entropy, styx_cwt, styx_fft, styx_stx, tfr_info  # noqa: E402
from quantum_inferno_b200.utilities import short_time_fft as stf  # noqa: E402

FS = 800.0
rng = np.random.default_rng(0)
for logn, order, ch in ((13, 3, 3), (15, 6, 2), (14, 12, 1), (16, 1.5, 2)):
    x = rng.standard_normal((ch, 1 << logn)).astype(np.float32)
    r = cwt_entropy.cwt_power_entropy(order, torch.from_numpy(x).cuda(), FS, dtype="float32", method="multirate")
    f, t, c = styx_cwt.cwt_complex_any_scale_pow2(order, x[0], FS, dtype="float32", method="multirate")
    print("multirate", logn, order, float(r.entropy_bits()[0]), c.shape, flush=True)
x = rng.standard_normal(5000)
f, t, c = styx_cwt.cwt_complex_any_scale_pow2(3, x, FS)
f, t, s = styx_stx.stx_complex_any_scale_pow2(3, x[:4096], FS)
f, t, z = styx_fft.stft_complex_pow2(x, FS, 200, overlap_points=150, nfft_points=512)
c, cb, t, f = cwt_atoms.cwt_chirp_from_sig(x[:2048], FS)
sh = tfr_info.shannon_stft_from_tfr_power(np.abs(c) ** 2)
for m, ov, pad in ((256, 128, "zeros"), (200, 150, "even"), (100, 20, "odd")):
    f, t, mag = stf.stft_tukey(x, FS, 0.25, m, ov, padding=pad)
    obj = stf.get_stft_object_tukey(FS, 0.25, m, ov)
    ts, xr = stf.istft_tukey(obj.stft(x), FS, 0.25, m, ov)
    print("stft_tukey", m, ov, pad, mag.shape, float(np.abs(xr[:len(x)] - x[:len(xr)]).max()), flush=True)
torch.cuda.synchronize()
print("done")
