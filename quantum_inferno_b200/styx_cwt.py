"""
Gabor-atom continuous wavelet transform of order N on the B200 -- drop-in for ``quantum_inferno.styx_cwt``.

Same function names, positional/keyword signatures and return tuples as the reference
(quantum_inferno/styx_cwt.py); the atoms and the FFT convolution run in CUDA kernels (csrc/qi_cwt.cu).
numpy in -> numpy out (complex128 / float64 by default, like the reference); a CUDA ``torch.Tensor`` in ->
tensors out, left on the device.  Keyword-only extras that do not exist upstream:

    dtype     'float64' (default) or 'float32' -- arithmetic and output precision
    spectrum  'auto' | 'analytic' | 'table'    -- how each band's frequency response is obtained
    outputs   'complex' (default) | 'power' | 'both'
    method    'exact' (default: record FFT + per-band inverse FFT) | 'multirate' (float32 only, 2^m >= 8192 points:
              the decimation-pyramid fast path, complex TFR within ~3e-6 of the plane maximum)

2-D input [channels, points] is an extension: the reference applied to every row.
"""
from typing import Tuple, Union

import numpy as np

from . import _driver, _plan
from . import scales_dyadic as scales
from ._runtime import dtype_name, finish, get_runtime


def wavelet_variance_theory(amp: float, time_s: np.ndarray, scale: float, omega: float) -> Tuple[float, float]:
    """Nominal variance of the real and imaginary parts of a Gabor atom (reference styx_cwt.py:15-26)."""
    common = amp ** 2 / len(time_s) * 0.5 * np.sqrt(np.pi) * scale
    decay = np.exp(-(scale * omega) ** 2)
    return common / (1 + decay), common / (1 - decay)


def wavelet_amplitude(scale_atom: Union[np.ndarray, float]):
    """(amp_canonical, amp_unit_spectrum) (reference styx_cwt.py:29-40)."""
    return _plan.wavelet_amplitude(scale_atom)


def amplitude_convert_norm_to_spect(scale_atom: Union[np.ndarray, float]):
    """Ratio amp_unit_spectrum / amp_canonical (reference styx_cwt.py:43-55)."""
    amp_canonical, amp_unit_spectrum = _plan.wavelet_amplitude(scale_atom)
    return amp_unit_spectrum / amp_canonical


def wavelet_time(time_s: np.ndarray, offset_time_s: float, frequency_sample_rate_hz: float) -> np.ndarray:
    """Non-dimensional shifted time (reference styx_cwt.py:58-65)."""
    return frequency_sample_rate_hz * (time_s - offset_time_s)


def wavelet_complex(band_order_nth: float, time_s: np.ndarray, offset_time_s: float,
                    scale_frequency_center_hz: Union[np.ndarray, float], frequency_sample_rate_hz: float, *,
                    dtype=None):
    """Unit-modulus Gabor atoms exp(-x^2/2s^2) exp(i omega x) on an arbitrary time axis
    (reference styx_cwt.py:68-110).  Returns (wavelet_gabor, xtime_shifted, scale_angular_frequency, scale,
    omega, amp_canonical, amp_unit_spectrum) with the reference's tiled [band x time] shapes."""
    rt = get_runtime()
    dt = dtype_name(dtype)
    xtime = wavelet_time(np.asarray(time_s, dtype=np.float64), offset_time_s, frequency_sample_rate_hz)
    scalar = np.isscalar(scale_frequency_center_hz)
    bands, scale_atom, omega_atom, _ = _plan.gabor_bands(band_order_nth, len(xtime), scale_frequency_center_hz,
                                                         frequency_sample_rate_hz, "unit", dt, "table")
    atoms = rt.to_numpy(_driver.atoms_time(bands, len(xtime), frequency_sample_rate_hz, dt, xtime=xtime, rt=rt))
    if scalar:
        scale, omega, atoms = float(scale_atom[0]), float(omega_atom[0]), atoms[0]
        angular = omega
    else:
        scale = np.tile(scale_atom, (len(xtime), 1)).T
        omega = np.tile(omega_atom, (len(xtime), 1)).T
        angular = omega_atom
    amp_canonical, amp_unit_spectrum = _plan.wavelet_amplitude(scale)
    return atoms, xtime, angular, scale, omega, amp_canonical, amp_unit_spectrum


def wavelet_centered_4cwt(band_order_nth: float, duration_points: int,
                          scale_frequency_center_hz: Union[np.ndarray, float], frequency_sample_rate_hz: float,
                          dictionary_type: str = "norm", *, dtype=None):
    """Gabor atoms centred on a record of ``duration_points`` samples (reference styx_cwt.py:113-144).
    dictionary_type: 'norm' (canonical unit norm), 'spect' (unit spectrum), 'unit' (unit modulus);
    anything else behaves as 'norm', as upstream.  Returns (atoms, time_s, scale, omega, amp)."""
    rt = get_runtime()
    dt = dtype_name(dtype)
    scalar = np.isscalar(scale_frequency_center_hz)
    bands, scale_atom, omega_atom, amp_atom = _plan.gabor_bands(
        band_order_nth, duration_points, scale_frequency_center_hz, frequency_sample_rate_hz, dictionary_type, dt, "table")
    atoms = rt.to_numpy(_driver.atoms_time(bands, duration_points, frequency_sample_rate_hz, dt, rt=rt))
    time_s = np.arange(duration_points) / frequency_sample_rate_hz
    xtime = frequency_sample_rate_hz * (time_s - time_s[-1] / 2.)
    if scalar:
        return atoms[0], xtime / frequency_sample_rate_hz, float(scale_atom[0]), float(omega_atom[0]), float(amp_atom[0])
    tile = lambda v: np.tile(v, (duration_points, 1)).T          # noqa: E731
    return atoms, xtime / frequency_sample_rate_hz, tile(scale_atom), tile(omega_atom), tile(amp_atom)


def cwt_complex_any_scale_pow2(band_order_nth: float, sig_wf, frequency_sample_rate_hz: float,
                               cwt_type: str = "fft", dictionary_type: str = "norm", *,
                               dtype=None, spectrum: str = "auto", outputs: str = "complex",
                               method: str = "exact"):
    """CWT of ``sig_wf`` with the order-N Gabor dictionary (reference styx_cwt.py:147-198):
    linear 'same' convolution with every atom, band centres from
    ``scales_dyadic.log_frequency_hz_from_fft_points`` (base G3).

    :return: frequency_cwt_hz [B], time_cwt_s [N], cwt [B, N] (or [C, B, N] for 2-D input)
    """
    if cwt_type == "morlet2":
        # upstream calls scipy.signal.cwt here, which SciPy >= 1.15 no longer has (AttributeError)
        raise NotImplementedError("cwt_type='morlet2' relied on scipy.signal.cwt, removed from SciPy; use 'fft'")
    rt = get_runtime(sig_wf)
    dt = dtype_name(dtype)
    want_numpy = not rt.is_device_array(sig_wf)
    sig, was_1d = _driver._as_2d(rt, sig_wf, dt)
    n_points = int(sig.shape[1])
    frequency_cwt_hz = scales.log_frequency_hz_from_fft_points(
        frequency_sample_hz=frequency_sample_rate_hz, fft_points=n_points, scale_order=band_order_nth)
    time_cwt_s = np.arange(n_points) / frequency_sample_rate_hz
    bands, scale, _, _ = _plan.gabor_bands(band_order_nth, n_points, frequency_cwt_hz, frequency_sample_rate_hz,
                                           dictionary_type, dt, spectrum)
    if method not in ("exact", "multirate"):
        raise ValueError("method must be 'exact' or 'multirate'")
    if method == "multirate":
        if dt != "float32" or not _plan.multirate_supported(n_points, scale):
            raise ValueError("method='multirate' needs dtype='float32' and a record of 2^m >= 8192 points")
        mr_bands, _, _, _ = _plan.multirate_bands(band_order_nth, n_points, frequency_cwt_hz,
                                                  frequency_sample_rate_hz, dictionary_type)
        res = _driver.cwt_multirate(sig, mr_bands, want_power=outputs in ("power", "both"),
                                    want_complex=outputs in ("complex", "both"), rt=rt)
    else:
        res = _driver.cwt_fft(sig, bands, frequency_sample_rate_hz, dt, want_complex=outputs in ("complex", "both"),
                              want_power=outputs in ("power", "both"), rt=rt)

    def shape(buf):
        if buf is None:
            return None
        return finish(rt, buf[0] if was_1d else buf, want_numpy)

    if outputs == "complex":
        return frequency_cwt_hz, time_cwt_s, shape(res["complex"])
    if outputs == "power":
        return frequency_cwt_hz, time_cwt_s, shape(res["power"])
    if outputs == "both":
        return frequency_cwt_hz, time_cwt_s, (shape(res["complex"]), shape(res["power"]))
    raise ValueError("outputs must be 'complex', 'power' or 'both'")
