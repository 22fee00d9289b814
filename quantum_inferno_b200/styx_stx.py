"""
Stockwell transform on the B200 -- drop-in for ``quantum_inferno.styx_stx`` (reference styx_stx.py).

The record FFT, the shifted-spectrum x Gaussian-window product and the per-band inverse FFTs run in
csrc/qi_stx.cu; the band table, the Gaussian widths and the integer shift indices are float64 host
arithmetic kept identical to the reference so the indices are bit-exact.  Records must have 2^m points.
Keyword-only extras: ``dtype`` ('float64' default / 'float32').  2-D input = one record per row.
"""
from typing import Tuple

import numpy as np

from . import _driver, _plan
from . import scales_dyadic as scales
from ._lib import STX_BAND
from ._runtime import dtype_name, finish, get_runtime
from .utilities.rescaling import is_power_of_two


def sig_pad_up_to_pow2(sig_wf: np.ndarray, n_fft: int, verbosity: bool = False):
    """Zero-pad to ``n_fft`` points (reference styx_stx.py:16-48), including its failure modes:
    n_fft=None raises TypeError at the first comparison and any real padding raises TypeError at the
    tuple + int concatenation, exactly as upstream."""
    n_times = sig_wf.shape[-1]
    if verbosity:
        print(f"length of fft: {n_fft}, waveform shape: {n_times}")
    if n_fft < n_times:                      # TypeError when n_fft is None, as upstream
        raise ValueError(f"n_fft cannot be smaller than signal size. Got {n_fft} < {n_times}.")
    if n_fft is None or (not is_power_of_two(n_fft) and n_times > n_fft):
        n_fft = 2 ** int(np.ceil(np.log2(n_times)))
    if n_times < n_fft:
        if verbosity:
            print(f'The input signal is shorter ({sig_wf.shape[-1]}) than "n_fft" ({n_fft}). Applying zero padding.')
        zero_pad: int = n_fft - n_times
        sig_wf = np.concatenate((sig_wf, np.zeros(sig_wf.shape[:-1] + zero_pad, sig_wf.dtype)), axis=-1)
    else:
        zero_pad: int = 0
    return sig_wf, n_fft, zero_pad


def _require_pow2(n_points):
    if not is_power_of_two(int(n_points)):
        raise ValueError(f"the Stockwell kernels need a record of 2^m points, got {n_points}")


def tfr_stx_fft(sig_wf, time_sample_interval: float, scale_order_input: float = 8.0, n_fft_in: int = None,
                frequency_min: float = None, frequency_max: float = None, frequency_step: float = None,
                factor_q: float = 0.0, power_p: float = 0.0, power_r: float = 1.0, is_geometric: bool = False,
                is_inferno: bool = False, scale_base_input: float = scales.Slice.G3,
                scale_ref_input: float = scales.Slice.T1S, *, dtype=None):
    """General Stockwell transform with linear / geometric / standardised frequency grids and the
    sigma-scaling exponents of Moukadem et al. (reference styx_stx.py:52-192).

    :return: tfr_stx, psd_stx, frequency_stx, frequency_stx_fft, windows_fft
    """
    rt = get_runtime(sig_wf)
    dt = dtype_name(dtype)
    want_numpy = not rt.is_device_array(sig_wf)
    host_sig = rt.to_numpy(sig_wf) if rt.is_device_array(sig_wf) else np.asarray(sig_wf)
    frequency_sample_rate: float = 1 / time_sample_interval
    cycles_m: float = 12.0 / 5.0 * scale_order_input
    sig_pow2, n_fft_pow2, zero_pad = sig_pad_up_to_pow2(host_sig, n_fft_in)
    _require_pow2(n_fft_pow2)
    n_fft_out = n_fft_pow2 - zero_pad

    frequency_fft = np.fft.fftfreq(n_fft_pow2, time_sample_interval)
    if frequency_min is None:
        frequency_min = cycles_m / (n_fft_pow2 / frequency_sample_rate)
    if frequency_max is None:
        frequency_max = frequency_sample_rate / 2.0
    f_start = frequency_fft[np.abs(frequency_fft - frequency_min).argmin()]
    f_stop = frequency_fft[np.abs(frequency_fft - frequency_max).argmin()]
    if frequency_step is None:
        frequency_step = (frequency_max - frequency_min) * 2.0 / len(frequency_fft)
    frequency_stx = np.arange(f_start, f_stop, frequency_step)
    if is_geometric is True:
        if is_inferno is True:
            frequency_stx = scales.band_frequency_low_high(
                frequency_order_input=scale_order_input, frequency_low_input=f_start, frequency_high_input=f_stop,
                frequency_sample_rate_input=frequency_sample_rate, frequency_base_input=scale_base_input,
                frequency_ref_input=scale_ref_input)[5]
        else:
            num_bands = int(np.log2(f_stop / f_start) * scale_order_input)
            frequency_stx = np.logspace(np.log2(f_start), np.log2(f_stop), num=num_bands, base=scale_base_input)

    bands = np.zeros(len(frequency_stx), dtype=STX_BAND)
    frequency_stx_fft = np.empty(len(frequency_stx))
    for isx, fsx in enumerate(frequency_stx):
        stx_index = _plan.nearest_fft_bin(fsx, n_fft_pow2, time_sample_interval)
        frequency_stx_fft[isx] = frequency_fft[stx_index]
        omega_sx = 2 * np.pi * frequency_stx_fft[isx] / frequency_sample_rate
        if omega_sx == 0.0:
            raise TypeError("object of type 'int' has no len()")       # upstream styx_stx.py:173
        sigma_scaling = (1 + factor_q * (omega_sx ** power_p)) * (omega_sx ** (1 - power_r))
        bands["sigma"][isx] = cycles_m / omega_sx * sigma_scaling
        bands["shift"][isx] = stx_index

    sig = rt.reshape(rt.asarray(sig_pow2, dt), (1, n_fft_pow2))
    res = _driver.stx_fft(sig, bands, dt, want_complex=True, rt=rt)
    tfr = res["complex"][0][:, :n_fft_out]
    if zero_pad > 0:
        tfr = tfr.contiguous() if hasattr(tfr, "contiguous") else np.ascontiguousarray(tfr)
    psd = _driver.abs_log2(tfr, dt, True, eps=scales.get_epsilon(), square=True, rt=rt)
    windows = _driver.stx_windows(bands, n_fft_pow2, dt, rt=rt)
    return (finish(rt, tfr, want_numpy), finish(rt, psd, want_numpy), frequency_stx, frequency_stx_fft,
            finish(rt, windows, want_numpy))


def stx_complex_any_scale_pow2(band_order_nth: float, sig_wf, frequency_sample_rate_hz: float, *,
                               dtype=None, outputs: str = "complex", method: str = "exact"):
    """Stockwell transform on the standard order-N band table (reference styx_stx.py:195-236).

    ``method="multirate"`` (keyword-only extension; float32, 2^m >= 4096 samples) computes every voice at its own
    decimated rate and interpolates it to the full rate: relative L2 error ~3e-6 against the exact method (north-star
    float32 tolerance 1e-4), a fraction of the memory traffic.

    :return: frequency_stx_hz [B], time_stx_s [N], tfr_stx [B, N] (or [C, B, N] for 2-D input)
    """
    rt = get_runtime(sig_wf)
    dt = dtype_name(dtype)
    want_numpy = not rt.is_device_array(sig_wf)
    sig, was_1d = _driver._as_2d(rt, sig_wf, dt)
    n_fft_pow2 = int(sig.shape[1])
    _require_pow2(n_fft_pow2)
    frequency_stx_hz, bands = _plan.stx_bands(band_order_nth, n_fft_pow2, frequency_sample_rate_hz)
    res = _driver.stx_fft(sig, bands, dt, want_complex=outputs != "power", want_power=outputs == "power", rt=rt,
                          method=method)
    buf = res["power"] if outputs == "power" else res["complex"]
    return (frequency_stx_hz, np.arange(n_fft_pow2) / frequency_sample_rate_hz,
            finish(rt, buf[0] if was_1d else buf, want_numpy))
