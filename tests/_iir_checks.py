"""Shared assertions for the zero-phase IIR drop-ins: run through the CPU emulator of the kernels
(tests/test_emul_api.py) and on the B200 (tests/test_gpu_parity.py, -m gpu).

Tolerances.  The reference's direct-form recursion is ill conditioned for its own band-passes: scipy's float64 result
differs from the same recursion in 80-bit arithmetic (`*_truth`, oracle/qi_oracle.py) by up to 3e-8 of the peak.  The
kernels evaluate the same transfer function as a cascade, so the bar is: within 1e-10 of the long-double arbiter, and
no farther from the reference than the reference is from the arbiter."""
import numpy as np
import pytest

FS = 800.0
BA_CALLS = {
    "lowpass_100": lambda m, x: m.butter_lowpass(x, FS, 100.0),
    "highpass_5": lambda m, x: m.butter_highpass(x, FS, 5.0),
    "bandpass_10_100": lambda m, x: m.butter_bandpass(x, FS, 10.0, 100.0),
    "bandpass_58_62": lambda m, x: m.butter_bandpass(x, FS, 58.0, 62.0, 4, 0.1),
    "bandpass_nyq": lambda m, x: m.butter_bandpass(x, FS, 20.0, 500.0, 3, 1.0),
    "lowpass_o5": lambda m, x: m.butter_lowpass(x, FS, 120.0, 5, 0.0),
}
SOS_CASES = {"pick_100_200_o7": ((100.0, 200.0), 7), "pick_1_10_o4": ((1.0, 10.0), 4), "pick_50_70_o2": ((50.0, 70.0), 2)}
TOL_TRUTH = 1e-10


def peak_rel(a, b):
    return float(np.max(np.abs(np.asarray(a, dtype=np.float64) - b)) / np.max(np.abs(b)))


def _judge(got, ref, truth, tag):
    assert got.shape == ref.shape and got.dtype == ref.dtype, tag
    e_truth, e_ref, ref_own = peak_rel(got, truth), peak_rel(got, ref), peak_rel(ref, truth)
    assert e_truth < TOL_TRUTH, (tag, e_truth)
    assert e_ref <= 2.0 * ref_own + 1e-11, (tag, e_ref, ref_own)


def check_butter_golden(golden, capsys):
    from quantum_inferno_b200 import styx_fft
    from quantum_inferno_b200.synth import synthetic_signals
    g = golden("iir")
    x = g["x"]
    for name, call in BA_CALLS.items():
        _judge(call(styx_fft, x), g[f"{name}_ref"], g[f"{name}_truth"], name)
    assert "greater than Nyquist" in capsys.readouterr().out                    # bandpass_nyq warns like the reference
    _judge(synthetic_signals.antialias_half_nyquist(x), g["antialias_ref"], g["antialias_truth"], "antialias")
    assert np.array_equal(synthetic_signals.taper_tukey(x[:100], 0.3), __import__("scipy.signal").signal.windows.tukey(100, 0.3))
    # batch axis and float32 records (fp64 arithmetic inside, float32 in / out)
    both = styx_fft.butter_bandpass(np.stack([x, x[::-1]]), FS, 10.0, 100.0)
    assert both.shape == (2, len(x)) and peak_rel(both[0], g["bandpass_10_100_truth"]) < TOL_TRUTH
    assert peak_rel(both[1], styx_fft.butter_bandpass(x[::-1].copy(), FS, 10.0, 100.0)) < 1e-14
    y32 = styx_fft.butter_lowpass(x.astype(np.float32), FS, 100.0)
    assert y32.dtype == np.float32 and peak_rel(y32, g["lowpass_100_truth"]) < 3e-7
    with pytest.raises(ValueError):
        styx_fft.butter_highpass(x, FS, 400.0)
    with pytest.raises(ValueError):
        styx_fft.butter_lowpass(x, FS, 401.0)
    with pytest.raises(ValueError):
        styx_fft.butter_lowpass(x[:15], FS, 100.0)                              # not longer than padlen = 15


def check_picker_bandpass_golden(golden):
    from quantum_inferno_b200.utilities import picker
    g = golden("iir")
    x = g["x"]
    for name, (band, order) in SOS_CASES.items():
        _judge(picker.apply_bandpass(x, band, FS, order), g[f"{name}_ref"], g[f"{name}_truth"], name)
    assert np.array_equal(picker.find_peaks_by_extraction_type_with_bandpass(x, (50.0, 70.0), FS, 4, "sigmax", 0.6),
                          g["peaks_bandpass"])
    # the reference's own error tests (quantum_inferno/tests/utilities/test_picker.py:56-59)
    for band in ((300, 100), (100, 100), (1, FS), (-1, 1)):
        with pytest.raises(ValueError):
            picker.apply_bandpass(x, band, FS)


def check_filtfilt_vs_oracle(n=9000, direct_form=True):
    """Ragged lengths around the 4096-sample tiles, both C-ABI forms, against the oracle's long-double recursion."""
    from scipy import signal
    from oracle import qi_oracle as orc
    from quantum_inferno_b200 import _driver, _iir, styx_fft
    from quantum_inferno_b200._runtime import get_runtime
    rt = get_runtime()
    rng = np.random.default_rng(8)
    b, a = signal.butter(4, [0.05, 0.4], btype="bandpass")
    for m in (28, 4096 - 54, 4096 - 53, 2 * 4096 - 54 + 1, n):
        x = rng.standard_normal(m) + 2.0
        truth = orc.butter_filtfilt(b, a, x, 0.25, dtype=np.longdouble).astype(np.float64)
        assert peak_rel(styx_fft._butter_filtfilt(x, b, a, 0.25), truth) < TOL_TRUTH, m
    # the direct-form entry of the C ABI (well-conditioned low-pass: same arithmetic as scipy's lfilter)
    if direct_form:
        b, a = signal.butter(4, 0.5)
        x = rng.standard_normal((3, n))
        got = rt.to_numpy(_driver.filtfilt(rt.asarray(x, "float64"), "float64", 15, b=b, a=a, zi=orc.lfilter_zi(b, a), rt=rt))
        for c in range(3):
            assert peak_rel(got[c], orc.filtfilt(b, a, x[c])) < 1e-13
    # exact re-factoring: the cascade is the same transfer function as the rounded taps
    for bb, aa in (signal.butter(4, [58 / 400, 62 / 400], btype="bandpass"), signal.butter(5, 0.3), signal.butter(1, 0.2, btype="highpass")):
        sos = _iir.tf2sos_exact(bb, aa)
        imp = np.zeros(400)
        imp[0] = 1.0
        want = orc.lfilter(bb, aa, imp.astype(np.longdouble), np.zeros(len(aa) - 1), dtype=np.longdouble)
        have = orc.sosfilt(sos, imp.astype(np.longdouble), np.zeros((sos.shape[0], 2)), dtype=np.longdouble)
        assert float(np.max(np.abs(want - have)) / np.max(np.abs(want))) < 1e-11
