"""
Deterministic benchmark signals made on the B200 -- drop-in for ``quantum_inferno.synth.benchmark_signals``
``well_tempered_tone`` (reference synth/benchmark_signals.py:268-355) and ``quantum_chirp`` (:57-109): same
arguments, printed warnings and return tuples.

The waveforms are synthesised where they are consumed (csrc/qi_synth.cu; float64 arithmetic in the reference's order
of operations), the anti-alias decimation of ``quantum_chirp`` is ``scipy.signal.decimate``'s Chebyshev cascade run by
csrc/qi_iir.cu followed by the strided copy of csrc/qi_pick.cu.  Keyword-only extras: ``dtype`` ("float64" default)
and ``device_out=True`` to get CUDA tensors instead of numpy arrays.  The generators of the reference that add
unseeded ``numpy.random`` noise (``synth_00`` .. ``synth_03``, ``add_noise_taper_aa=True``) are not reproducible and
are not provided.
"""
from typing import Tuple

import numpy as np

from .. import _driver
from .._runtime import dtype_name, get_runtime
from ..utilities.sampling import _decimate_device

DEFAULT_TIME_SAMPLE_INTERVAL = 1e-3
DEFAULT_TIME_DURATION = 1.0
DEFAULT_OVERSAMPLE_SCALE = 2


def quantum_chirp(omega: float, order: float = 12.0, gamma: float = 0.0, gauss: bool = True,
                  oversample_scale: int = DEFAULT_OVERSAMPLE_SCALE, *, dtype=None, device_out: bool = False
                  ) -> Tuple[np.ndarray, int]:
    """
    Gabor atom / sweep of 2^n points with a Gaussian window option, oversampled and anti-alias decimated
    (reference synth/benchmark_signals.py:57-109).

    :return: complex waveform, number of points of the window support (a power of two)
    """
    if omega >= 0.8 * np.pi:
        print("Omega >= 0.8*pi (AA*Nyquist), reset to pi * 2**(-1/N")
        omega = np.pi * 2 ** (-1 / order)
    scale_multiplier = 3.0 / 4.0 * np.pi * order
    scale = scale_multiplier / omega
    chirp_scale = scale * np.sqrt(1 + gamma ** 2)
    window_support_points = 2.0 * np.pi * chirp_scale
    window_support_pow2 = 2 ** int((np.ceil(np.log2(window_support_points))))
    n_over = oversample_scale * window_support_pow2
    rt = get_runtime()
    dt = dtype_name(dtype)
    # time = arange(n) - (n - 1) / 2  (the reference's time0 - time0[-1] / 2)
    both = _driver.synth_chirp(n_over, "float64", omega, t_center=float(n_over - 1) / 2, half_gamma=0.5 * gamma,
                               chirp_scale=chirp_scale, gauss=gauss, want_imag=True, rt=rt)
    dec = _decimate_device(rt, both, "float64", oversample_scale)            # real and imaginary part in one call
    if device_out:
        torch = rt.torch
        wf = torch.complex(dec[0], dec[1])
        return (wf if dt == "float64" else wf.to(torch.complex64)), window_support_pow2
    dec = rt.to_numpy(dec)
    wf = dec[0] + 1j * dec[1]
    return (wf if dt == "float64" else wf.astype(np.complex64)), window_support_pow2


def well_tempered_tone(frequency_sample_rate_hz: float = 800.0, frequency_center_hz: float = 60.0,
                       time_duration_s: float = 10.24, time_fft_s: float = 0.64, use_fft_frequency: bool = True,
                       add_noise_taper_aa: bool = False, output_desc: bool = False, *, dtype=None,
                       device_out: bool = False
                       ) -> Tuple[np.ndarray, np.ndarray, int, float, float, float]:
    """
    Tone of unit amplitude whose frequency sits on an FFT bin, 2^n points (reference
    synth/benchmark_signals.py:268-355).

    :return: waveform, timestamps, fft duration in points, sample rate, centre frequency, frequency resolution
    """
    if add_noise_taper_aa:
        raise NotImplementedError("add_noise_taper_aa draws unseeded numpy.random noise in the reference "
                                  "(synth/synthetic_signals.py:169); only the deterministic tone is provided")
    fs = frequency_sample_rate_hz

    def floor_pow2(seconds):
        """Largest power of two not above seconds * fs (the reference truncates log2 with int())."""
        return 2 ** (int(np.log2(seconds * fs)))

    n_record, n_fft = floor_pow2(time_duration_s), floor_pow2(time_fft_s)
    for what, asked, got in (("The time duration", time_duration_s, n_record), ("fft duration", time_fft_s, n_fft)):
        if got != asked * fs:                                           # same wording as the reference's two warnings
            print(f"Warning: {what} {asked} s with given sample rate doesn't produce data points "
                  f"that are power of two, adjusting {what[4:] if what.startswith('The ') else what} to {got} s")
    # the tone sits on a bin of the n_fft-point transform unless the caller asks for the nominal frequency
    bins_hz = np.fft.rfftfreq(n_fft, d=1 / fs)
    frequency_center_fft_hz = bins_hz[np.argmin(np.abs(bins_hz - frequency_center_hz))]
    frequency_resolution_fft_hz = fs / n_fft
    f_c = (frequency_center_fft_hz if use_fft_frequency else frequency_center_hz) / fs
    rt = get_runtime()
    dt = dtype_name(dtype)
    # mic_sig = cos(2.0 * pi * f_c * time_nd): the scalar product first, as Python evaluates it
    sig = rt.reshape(_driver.synth_chirp(n_record, dt, 2.0 * np.pi * f_c, rt=rt), (n_record,))
    time_s = np.arange(n_record) / fs
    if output_desc:
        print("WELL TEMPERED TONE SYNTHETIC")
        for label, value in (("Nyquist frequency:", fs / 2), ("Nominal signal frequency, hz:", frequency_center_hz),
                             ("FFT signal frequency, hz:", frequency_center_fft_hz),
                             ("Nominal spectral resolution, hz", 1.0 / time_fft_s),
                             ("FFT spectral resolution, hz", frequency_resolution_fft_hz),
                             ("Number of signal points:", n_record), ("log2(points):", np.log2(n_record)),
                             ("Number of FFT points:", n_fft), ("log2(FFT points):", np.log2(n_fft))):
            print(label, value)
    time_fft_nd = n_fft
    mic_sig = sig if device_out else rt.to_numpy(sig)
    return mic_sig, time_s, time_fft_nd, frequency_sample_rate_hz, frequency_center_fft_hz, frequency_resolution_fft_hz
