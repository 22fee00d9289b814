"""TEST INFRASTRUCTURE: every kernel family once, small sizes, for tests/emul/asan_check.sh (AddressSanitizer build of
the CPU kernel-logic emulator: out-of-bounds accesses to global buffers and shared memory)."""
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..")))
import numpy as np
from tests.emul.emul_runtime import EmulRuntime  # noqa: E402
from quantum_inferno_b200 import _runtime, cwt_entropy, styx_cwt, styx_stx, styx_fft, cwt_atoms
from quantum_inferno_b200.utilities import short_time_fft as stf
rng=np.random.default_rng(0)
with _runtime.use_runtime(EmulRuntime()):
    for logn, order, ch in ((13, 3, 2), (14, 6, 1), (13, 12, 1)):
        x = rng.standard_normal((ch, 1 << logn)).astype(np.float32)
        r = cwt_entropy.cwt_power_entropy(order, x, 800.0, dtype="float32", method="multirate")
        print("fused", logn, order, float(np.asarray(r.entropy_bits())[0]), flush=True)
    f,t,c = styx_cwt.cwt_complex_any_scale_pow2(3, x[0], 800.0, dtype="float32", method="multirate")
    print("complex multirate", c.shape, flush=True)
    x = rng.standard_normal(3000)
    f,t,c = styx_cwt.cwt_complex_any_scale_pow2(3, x, 800.0); print("exact", c.shape, flush=True)
    f,t,s = styx_stx.stx_complex_any_scale_pow2(3, x[:2048], 800.0); print("stx", s.shape, flush=True)
    f,t,z = styx_fft.stft_complex_pow2(x, 800.0, 200, overlap_points=150, nfft_points=512); print("stft", z.shape, flush=True)
    for m, ov, pad in ((256,128,"zeros"),(100,20,"odd")):
        f,t,mag = stf.stft_tukey(x, 800.0, 0.25, m, ov, padding=pad)
        o = stf.get_stft_object_tukey(800.0, 0.25, m, ov)
        ts, xr = stf.istft_tukey(o.stft(x), 800.0, 0.25, m, ov); print("tukey", mag.shape, xr.shape, flush=True)
    # round 2: band-limited routes of the exact paths (overlap-save blocks, decimated transforms + interpolation,
    # factorised inter-pass twiddles), both dtypes; Bluestein rfft; fused Stockwell entropy
    from quantum_inferno_b200 import tfr_info
    x = rng.standard_normal((2, 1 << 13))
    for dt in ("float64", "float32"):
        f,t,c = styx_cwt.cwt_complex_any_scale_pow2(3, x, 800.0, dtype=dt); print("exact band-limited", dt, c.shape, flush=True)
        f,t,s = styx_stx.stx_complex_any_scale_pow2(3, x, 800.0, dtype=dt); print("stx band-limited", dt, s.shape, flush=True)
        r = cwt_entropy.stx_power_entropy(3, x, 800.0, dtype=dt); print("stx entropy", dt, float(np.asarray(r.entropy_bits())[0]), flush=True)
    ff = tfr_info.ShannonFFT(x[0, :777]); print("bluestein", ff.sig.shape, flush=True)
print("asan run done")
