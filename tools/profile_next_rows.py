"""One call of each kernel family of the SURVEY 8(f) rows 3 / 4 at bench size, for `ncu --set full` captures
(profiles/r01_next_rows_*).  Not a benchmark: numbers printed under a profiler are never quoted."""
import os
import sys

import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from bench import FS, synth_batch_torch  # noqa: E402
from quantum_inferno_b200 import styx_fft  # noqa: E402
from quantum_inferno_b200.utilities import short_time_fft as stf  # noqa: E402
from quantum_inferno_b200.utilities import sampling  # noqa: E402

DEV = torch.device("cuda", 0)
which = sys.argv[1:] or ["sub", "iir", "stft"]
if "sub" in which:
    p = torch.rand((60, 1 << 24), device=DEV, dtype=torch.float32)
    for m in ("average", "max", "median"):
        sampling.subsample_2d(p, 32, m)
        sampling.subsample_2d(p, 2048, m)
    del p
if "iir" in which:
    x = synth_batch_torch(torch, 1 << 22, list(range(16)), DEV).double()
    styx_fft.butter_bandpass(x, FS, 10.0, 100.0)
if "stft" in which:
    x = synth_batch_torch(torch, 1 << 20, list(range(64)), DEV)
    styx_fft.stft_complex_pow2(x, FS, 1024, alpha=1.0, dtype="float32")
    obj = stf.get_stft_object_tukey(FS, 0.25, 1024, 512, dtype="float32")
    stf.istft_tukey(obj.stft(x), FS, 0.25, 1024, 512, dtype="float32")
if "stxmr" in which:
    from quantum_inferno_b200 import styx_stx
    x = synth_batch_torch(torch, 1 << 18, list(range(16)), DEV)
    styx_stx.stx_complex_any_scale_pow2(3, x, FS, dtype="float32", method="multirate")
torch.cuda.synchronize()
print("done")
