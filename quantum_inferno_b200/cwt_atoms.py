"""
Quantised Gabor "chirp" atoms and their CWT on the B200 -- drop-in for ``quantum_inferno.cwt_atoms``
(reference cwt_atoms.py): exact-Q cycles M = 2*Q*gamma, true base-2 band tables, optional sweep index.

The per-band Python loop of the reference (atom -> fft -> product -> ifft, cwt_atoms.py:406-435) becomes one
batched launch sequence in csrc/qi_cwt.cu: the complex-Gaussian atoms are synthesised on the device, transformed
once, and multiplied against the record spectrum inside the inverse-FFT kernel.  Scalar helpers stay on the host
in float64.  "fft" (N-point circular correlation, N = 2^m) and "conv" (linear, 'same') are both provided;
"morlet2" depended on scipy.signal.cwt, which SciPy removed, and raises NotImplementedError.
"""
from typing import Tuple, Union

import numpy as np

from . import _driver
from . import scales_dyadic as scales
from ._lib import ATOM_BAND, QI_CONV_CIRC_CORR, QI_CONV_LINEAR_SAME
from ._runtime import dtype_name, finish, get_runtime
from .utilities.rescaling import is_power_of_two


def chirp_mqg_from_n(band_order_nth: float, index_shift: float = 0, scale_base: float = scales.Slice.G2
                     ) -> Tuple[float, float, float]:
    """(cycles M, quality factor Q, gamma) for order N (reference cwt_atoms.py:122-144)."""
    if band_order_nth < 0.7:
        band_order_nth = 3.0
        print(f"N < 0.7 specified, using N = {band_order_nth}")
    band_edge = scale_base ** (1.0 / 2.0 / band_order_nth)
    quality_factor_q = 1.0 / (band_edge - 1.0 / band_edge)
    gamma = np.sqrt(np.log(2)) * (1 - np.log(2) * (index_shift / np.pi) ** 2) ** (-0.5)
    return 2 * quality_factor_q * gamma, quality_factor_q, gamma


def chirp_spectrum(frequency_hz: np.ndarray, offset_time_s: float, band_order_nth: float, frequency_center_hz: float,
                   frequency_sample_rate_hz: float, index_shift: float = 0, scale_base: float = scales.Slice.G2):
    """Closed-form spectrum of the quantum chirp on a frequency axis (reference cwt_atoms.py:53-92; host float64).

    :return: spectrum (complex), frequency shifted by the band centre [Hz]
    """
    cycles_m, _, gamma = chirp_mqg_from_n(band_order_nth, index_shift, scale_base)
    scale_atom = chirp_scale(cycles_m, frequency_center_hz, frequency_sample_rate_hz)
    p_complex = chirp_p_complex(scale_atom, gamma, index_shift)
    omega_centre = 2 * np.pi * frequency_center_hz / frequency_sample_rate_hz
    omega = 2 * np.pi * frequency_hz / frequency_sample_rate_hz
    omega_shifted = omega - omega_centre
    envelope = np.exp(-(1.0 / (4 * p_complex)) * (omega_shifted ** 2))
    spectrum = np.sqrt(p_complex / np.abs(p_complex)) * envelope * np.exp(-1j * 2 * np.pi * frequency_hz * offset_time_s)
    return spectrum, omega_shifted * frequency_sample_rate_hz / (2 * np.pi)


def chirp_spectrum_centered(band_order_nth: float, scale_frequency_center_hz: float, frequency_sample_rate_hz: float,
                            index_shift: float = 0, scale_base: float = scales.Slice.G2):
    """Same spectrum on the fixed grid [-pi, pi) in steps of pi/128 about the band centre
    (reference cwt_atoms.py:95-119).

    :return: spectrum (complex), shifted frequency [Hz]
    """
    cycles_m, _, gamma = chirp_mqg_from_n(band_order_nth, index_shift, scale_base)
    scale_atom = chirp_scale(cycles_m, scale_frequency_center_hz, frequency_sample_rate_hz)
    p_complex = chirp_p_complex(scale_atom, gamma, index_shift)
    omega_shifted = np.arange(-np.pi, np.pi, np.pi / 2 ** 7)
    spectrum = np.sqrt(p_complex / np.abs(p_complex)) * np.exp(-(omega_shifted ** 2) / (4 * p_complex))
    return spectrum, omega_shifted * frequency_sample_rate_hz / (2 * np.pi)


def chirp_scale(cycles_m: float, scale_frequency_center_hz: Union[np.ndarray, float],
                frequency_sample_rate_hz: float) -> float:
    """Non-dimensional atom scale M*fs/(2*pi*fc) (reference cwt_atoms.py:147-158)."""
    return cycles_m * frequency_sample_rate_hz / scale_frequency_center_hz / (2.0 * np.pi)


def chirp_scale_from_order(band_order_nth: float, scale_frequency_center_hz: float, frequency_sample_rate_hz: float,
                           index_shift: float = 0, scale_base: float = scales.Slice.G2) -> float:
    """Scale from the order; argument order of the inner call is the reference's (cwt_atoms.py:161-180)."""
    cycles_m, _, _ = chirp_mqg_from_n(band_order_nth, index_shift, scale_base)
    return chirp_scale(cycles_m, frequency_sample_rate_hz, scale_frequency_center_hz)


def chirp_uncertainty(scale_atom: float, frequency_sample_rate_hz: float, gamma: float, index_shift: float
                      ) -> Tuple[float, float, float]:
    """(time std [s], frequency std [Hz], angular frequency std [Hz]) (reference cwt_atoms.py:183-199)."""
    time_std_s = scale_atom / np.sqrt(2) / frequency_sample_rate_hz
    angular_frequency_std = np.sqrt(1 + (index_shift * gamma) ** 2) / scale_atom / np.sqrt(2)
    angular_frequency_std_hz = frequency_sample_rate_hz * angular_frequency_std
    return time_std_s, angular_frequency_std_hz / 2 / np.pi, angular_frequency_std_hz


def chirp_p_complex(scale_atom: float, gamma: float, index_shift: float) -> complex:
    """Complex Gaussian coefficient p (reference cwt_atoms.py:202-211)."""
    return (1 - 1j * index_shift * gamma / np.pi) / (2 * scale_atom ** 2)


def chirp_amplitude(scale_atom: float, gamma: float, index_shift: float) -> Tuple[float, float]:
    """(normal_scaling, spectrum_scaling) (reference cwt_atoms.py:214-226)."""
    p_complex = chirp_p_complex(scale_atom, gamma, index_shift)
    return 1 / np.pi ** 0.25 * 1 / np.sqrt(scale_atom), np.sqrt(np.abs(p_complex) / np.pi)


def chirp_time(time_s: np.ndarray, offset_time_s: float, frequency_sample_rate_hz: float) -> np.ndarray:
    """Scaled, shifted time (reference cwt_atoms.py:229-238)."""
    return frequency_sample_rate_hz * (time_s - offset_time_s)


def chirp_scales_from_duration(band_order_nth: float, sig_duration_s: float, index_shift: float = 0.0,
                               scale_base: float = scales.Slice.G2) -> Tuple[float, float]:
    """(time scale [s], frequency scale [Hz]) of the longest atom (reference cwt_atoms.py:241-256)."""
    cycles_m, _, _ = chirp_mqg_from_n(band_order_nth, index_shift, scale_base)
    scale_time_s = sig_duration_s / cycles_m
    return scale_time_s, 1 / scale_time_s


def chirp_frequency_bands(scale_order_input: float, frequency_low_input: float, frequency_sample_rate_input: float,
                          frequency_high_input: float, index_shift: float = 0,
                          frequency_ref: float = scales.Slice.F1HZ, scale_base: float = scales.Slice.G2):
    """(order, M, Q, gamma, band centres (descending), band starts, band ends) (reference cwt_atoms.py:259-300)."""
    order_nth, scale_base, _, frequency_ref, _, centre, start, end = scales.band_frequency_low_high(
        frequency_order_input=scale_order_input, frequency_base_input=scale_base, frequency_ref_input=frequency_ref,
        frequency_low_input=frequency_low_input, frequency_high_input=frequency_high_input,
        frequency_sample_rate_input=frequency_sample_rate_input)
    cycles_m, quality_q, gamma = chirp_mqg_from_n(order_nth, index_shift, scale_base)
    return order_nth, cycles_m, quality_q, gamma, centre, start, end


def _chirp_bands(band_order_nth, frequency_hz, fs, index_shift, scale_base, dictionary_type):
    """ATOM_BAND rows for chirp atoms: a(x) = A exp(-p x^2) exp(i M x / s)."""
    f = np.atleast_1d(np.asarray(frequency_hz, dtype=np.float64))
    cycles_m, _, gamma = chirp_mqg_from_n(band_order_nth, index_shift, scale_base)
    scale_atom = chirp_scale(cycles_m, f, fs)
    p_complex = chirp_p_complex(scale_atom, gamma, index_shift)
    normal_scaling, spectrum_scaling = chirp_amplitude(scale_atom, gamma, index_shift)
    bands = np.zeros(len(f), dtype=ATOM_BAND)
    bands["omega"] = cycles_m / scale_atom
    bands["p_re"] = np.real(p_complex)
    bands["p_im"] = np.imag(p_complex)
    bands["amp"] = normal_scaling if dictionary_type == "norm" else spectrum_scaling
    bands["analytic"] = 0
    return bands, normal_scaling, spectrum_scaling


def chirp_complex(band_order_nth: float, time_s: np.ndarray, offset_time_s: float, scale_frequency_center_hz: float,
                  frequency_sample_rate_hz: float, index_shift: float = 0, scale_base: float = scales.Slice.G2, *,
                  dtype=None):
    """Unscaled quantum chirp on an arbitrary time axis (reference cwt_atoms.py:16-50).

    :return: waveform_complex, time_shifted_s, normal_scaling, spectrum_scaling
    """
    rt = get_runtime()
    dt = dtype_name(dtype)
    xtime = chirp_time(np.asarray(time_s, dtype=np.float64), offset_time_s, frequency_sample_rate_hz)
    bands, normal_scaling, spectrum_scaling = _chirp_bands(band_order_nth, scale_frequency_center_hz,
                                                           frequency_sample_rate_hz, index_shift, scale_base, "norm")
    bands["amp"] = 1.0
    atom = rt.to_numpy(_driver.atoms_time(bands, len(xtime), frequency_sample_rate_hz, dt, xtime=xtime, rt=rt))[0]
    return atom, xtime / frequency_sample_rate_hz, float(normal_scaling[0]), float(spectrum_scaling[0])


def chirp_centered_4cwt(band_order_nth: float, sig_or_time: np.ndarray, scale_frequency_center_hz: float,
                        frequency_sample_rate_hz: float, index_shift: float = 0,
                        scale_base: float = scales.Slice.G2, dictionary_type: str = "norm", *, dtype=None):
    """Atom centred on a record as long as ``sig_or_time`` (reference cwt_atoms.py:303-340).

    :return: waveform_complex, time_shifted_s
    """
    rt = get_runtime()
    dt = dtype_name(dtype)
    n = len(sig_or_time)
    bands, _, _ = _chirp_bands(band_order_nth, scale_frequency_center_hz, frequency_sample_rate_hz, index_shift,
                               scale_base, dictionary_type)
    atom = rt.to_numpy(_driver.atoms_time(bands, n, frequency_sample_rate_hz, dt, rt=rt))[0]
    time_s = np.arange(n) / frequency_sample_rate_hz
    return atom, chirp_time(time_s, time_s[-1] / 2.0, frequency_sample_rate_hz) / frequency_sample_rate_hz


def cwt_chirp_complex(band_order_nth: float, sig_wf, frequency_low_hz: float, frequency_sample_rate_hz: float,
                      frequency_high_hz: float = scales.Slice.F0HZ, cwt_type: str = "fft", index_shift: float = 0,
                      frequency_ref: float = scales.Slice.F1HZ, scale_base: float = scales.Slice.G2,
                      dictionary_type: str = "norm", *, dtype=None):
    """CWT with the chirp dictionary (reference cwt_atoms.py:343-444).

    :return: cwt [B, N], cwt_bits [B, N], time_s [N], frequency_cwt_hz [B] (ascending)
    """
    if cwt_type == "morlet2":
        raise NotImplementedError("cwt_type='morlet2' relied on scipy.signal.cwt, removed from SciPy")
    if cwt_type not in ("fft", "conv"):
        raise ValueError(f"Incorrect cwt_type: {cwt_type} specified in cwt_chirp_complex")
    rt = get_runtime(sig_wf)
    dt = dtype_name(dtype)
    want_numpy = not rt.is_device_array(sig_wf)
    sig, was_1d = _driver._as_2d(rt, sig_wf, dt)
    wavelet_points = int(sig.shape[1])
    time_s = np.arange(wavelet_points) / frequency_sample_rate_hz
    if frequency_high_hz > frequency_sample_rate_hz / 2.0:
        frequency_high_hz = frequency_sample_rate_hz / 2.0
    order_nth, _, _, _, frequency_descending, _, _ = chirp_frequency_bands(
        scale_order_input=band_order_nth, frequency_low_input=frequency_low_hz,
        frequency_sample_rate_input=frequency_sample_rate_hz, frequency_high_input=frequency_high_hz,
        index_shift=index_shift, frequency_ref=frequency_ref, scale_base=scale_base)
    frequency_cwt_hz = np.flip(frequency_descending)
    # rows are produced directly in ascending-frequency order (the reference computes descending, then flips)
    bands, _, _ = _chirp_bands(order_nth, frequency_cwt_hz, frequency_sample_rate_hz, index_shift, scale_base,
                               dictionary_type)
    if cwt_type == "fft":
        if not is_power_of_two(wavelet_points):
            raise ValueError(f"cwt_type='fft' needs a record of 2^m points on this path, got {wavelet_points}")
        mode = QI_CONV_CIRC_CORR
    else:
        mode = QI_CONV_LINEAR_SAME
    res = _driver.cwt_fft(sig, bands, frequency_sample_rate_hz, dt, conv_mode=mode, want_complex=True, rt=rt)
    cwt = res["complex"]
    cwt_bits = _driver.abs_log2(cwt, dt, True, eps=scales.get_epsilon(), rt=rt)
    if was_1d:
        cwt, cwt_bits = cwt[0], cwt_bits[0]
    return finish(rt, cwt, want_numpy), finish(rt, cwt_bits, want_numpy), time_s, frequency_cwt_hz


def cwt_chirp_from_sig(sig_wf, frequency_sample_rate_hz: float, band_order_nth: float = 3, cwt_type: str = "fft",
                       index_shift: float = 0, frequency_ref: float = scales.Slice.F1HZ,
                       scale_base: float = scales.Slice.G2, dictionary_type: str = "norm", *, dtype=None):
    """CWT over every band that fits the record, up to Nyquist (reference cwt_atoms.py:447-486).

    :return: cwt, cwt_bits, time_s, frequency_cwt_hz
    """
    n_points = int(sig_wf.shape[-1])
    _, min_frequency_hz = chirp_scales_from_duration(
        band_order_nth=band_order_nth, sig_duration_s=n_points / frequency_sample_rate_hz, index_shift=index_shift,
        scale_base=scale_base)
    return cwt_chirp_complex(band_order_nth=band_order_nth, sig_wf=sig_wf, frequency_low_hz=min_frequency_hz,
                             frequency_sample_rate_hz=frequency_sample_rate_hz,
                             frequency_high_hz=frequency_sample_rate_hz / 2.0, cwt_type=cwt_type,
                             index_shift=index_shift, frequency_ref=frequency_ref, scale_base=scale_base,
                             dictionary_type=dictionary_type, dtype=dtype)
