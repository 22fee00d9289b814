"""
make_golden_iir.py -- generates tests/golden/iir.npz by running the UNMODIFIED reference
(quantum_inferno.styx_fft.butter_*, synth.synthetic_signals.antialias_half_nyquist, utilities.picker.apply_bandpass,
imported from /root/reference with the installed scipy behind them) on a seeded record, and the long-double
restatement of the same recursions (oracle/qi_oracle.py) as the arbiter.  SURVEY 8(f) rank 4.  Build container only.
    python oracle/make_golden_iir.py
"""
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
REF = os.environ.get("QI_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
from scipy import signal  # noqa: E402
from quantum_inferno import styx_fft  # noqa: E402
from quantum_inferno.synth import synthetic_signals  # noqa: E402
from quantum_inferno.utilities import picker  # noqa: E402
from oracle import qi_oracle as orc  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
FS = 800.0
# name, reference call, (b, a) the reference designs, tukey alpha
BA_CASES = [
    ("lowpass_100", lambda x: styx_fft.butter_lowpass(x, FS, 100.0), signal.butter(4, [100.0 / 400.0], btype="lowpass"), 0.5),
    ("highpass_5", lambda x: styx_fft.butter_highpass(x, FS, 5.0), signal.butter(4, [5.0 / 400.0], btype="highpass"), 0.5),
    ("bandpass_10_100", lambda x: styx_fft.butter_bandpass(x, FS, 10.0, 100.0), signal.butter(4, [0.025, 0.25], btype="bandpass"), 0.5),
    ("bandpass_58_62", lambda x: styx_fft.butter_bandpass(x, FS, 58.0, 62.0, 4, 0.1), signal.butter(4, [58.0 / 400.0, 62.0 / 400.0], btype="bandpass"), 0.1),
    ("bandpass_nyq", lambda x: styx_fft.butter_bandpass(x, FS, 20.0, 500.0, 3, 1.0), signal.butter(3, [0.05, 0.5], btype="bandpass"), 1.0),
    ("lowpass_o5", lambda x: styx_fft.butter_lowpass(x, FS, 120.0, 5, 0.0), signal.butter(5, [0.3], btype="lowpass"), 0.0),
    ("antialias", lambda x: synthetic_signals.antialias_half_nyquist(x), signal.butter(4, 0.5, btype="lowpass"), None),
]
SOS_CASES = [("pick_100_200_o7", (100.0, 200.0), 7), ("pick_1_10_o4", (1.0, 10.0), 4), ("pick_50_70_o2", (50.0, 70.0), 2)]


def main():
    n = 6000
    k = np.arange(n)
    x = np.random.default_rng(2024).standard_normal(n) + np.cos(2 * np.pi * 60.0 / FS * k) + 0.5 + 1e-3 * k
    d = {"x": x}
    for name, call, (b, a), alpha in BA_CASES:
        ref = call(x.copy())
        xt = x if alpha is None else x * signal.windows.tukey(n, alpha)
        assert np.array_equal(ref, signal.filtfilt(b, a, xt)), name           # the taps below ARE the reference's
        d[f"{name}_b"], d[f"{name}_a"], d[f"{name}_ref"] = b, a, ref
        d[f"{name}_alpha"] = np.array(-1.0 if alpha is None else alpha)
        d[f"{name}_truth"] = orc.filtfilt(b, a, xt.astype(np.longdouble), dtype=np.longdouble).astype(np.float64)
    for name, band, order in SOS_CASES:
        ref = picker.apply_bandpass(x, band, FS, order)
        sos = signal.butter(order, band, fs=FS, btype="band", output="sos")
        assert np.array_equal(ref, signal.sosfiltfilt(sos, x)), name
        d[f"{name}_sos"], d[f"{name}_ref"] = sos, ref
        d[f"{name}_truth"] = orc.sosfiltfilt(sos, x.astype(np.longdouble), dtype=np.longdouble).astype(np.float64)
    d["peaks_bandpass"] = picker.find_peaks_by_extraction_type_with_bandpass(x, (50.0, 70.0), FS, 4, "sigmax", 0.6)
    np.savez_compressed(os.path.join(OUT, "iir.npz"), **d)
    print("iir.npz", os.path.getsize(os.path.join(OUT, "iir.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
