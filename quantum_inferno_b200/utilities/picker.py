"""
Peak picking on the B200 -- drop-in for the scan functions of ``quantum_inferno.utilities.picker``
(reference utilities/picker.py:34-53 ``scale_signal_by_extraction_type``, :108-120
``find_peaks_by_extraction_type``, :123-149 ``find_peaks_with_bits``, :152-209 index helpers; same names,
arguments and return values).

The O(n) work runs in csrc/qi_pick.cu on the record where it lies (HBM): extrema (``qi_extrema``), log2 scaling
(``qi_abs_log2``), and the plateau-aware local-maximum scan of ``scipy.signal.find_peaks`` fused with its height
test (``qi_local_maxima``).  What is left for the host is the short list of candidate peaks: sorting it and
scipy's priority-ordered minimum-distance selection (``_select_by_peak_distance``), O(#peaks).  Peak indices are
bit-exact.  The Butterworth band-pass variants (:56-105) design the sections with scipy on the host and run
``sosfiltfilt`` over the record as the blocked parallel scan of csrc/qi_iir.cu.
"""
from typing import Optional, Tuple, Union

import numpy as np

from .. import _driver, _lib
from .._runtime import finish, get_runtime
from ..scales_dyadic import get_epsilon

INPUT_SCALE_TYPE = ["amplitude", "log2"]
EXTRACTION_TYPE = ["sigmax", "sigmin", "sigabs", "log2", "log2max"]


def _device_record(timeseries):
    rt = get_runtime(timeseries)
    if rt.is_device_array(timeseries):
        name = str(timeseries.dtype).replace("torch.", "")
    else:
        name = np.asarray(timeseries).dtype.name
    dt = name if name in ("float32", "float64") else "float64"
    x = rt.asarray(timeseries, dt)
    if x.ndim != 1:
        raise ValueError("`x` must be a 1-D array")           # scipy.signal.find_peaks' own message
    return rt, x, dt, not rt.is_device_array(timeseries)


def _extrema(rt, x, dt):
    """(nanmax, nanmin, nanmax|x|, has_nan) of a device record."""
    e = _driver.extrema(rt.reshape(x, (1, x.shape[0])), dt, rt=rt)[0]
    return float(e[0]), float(e[1]), float(e[2]), bool(e[3] > 0)


def _scaled(rt, x, dt, extraction_type):
    """Device record scaled as reference utilities/picker.py:34-53."""
    if extraction_type not in EXTRACTION_TYPE:
        print("Invalid extraction type.  Defaulting to sigmax.")
        extraction_type = "sigmax"
    if extraction_type in ("log2", "log2max"):
        bits = _driver.abs_log2(x, dt, False, eps=get_epsilon(), rt=rt)
        if extraction_type == "log2":
            return bits
        return _driver.divide(bits, dt, _extrema(rt, bits, dt)[0], rt=rt)
    vmax, vmin, vabs, _ = _extrema(rt, x, dt)
    return _driver.divide(x, dt, {"sigmax": vmax, "sigmin": vmin, "sigabs": vabs}[extraction_type], rt=rt)


def scale_signal_by_extraction_type(in_signal: np.ndarray, extraction_type: str = "sigmax") -> np.ndarray:
    """
    Normalise the signal by its (NaN-ignoring) maximum, minimum, absolute maximum, or convert it to log2 bits
    (reference utilities/picker.py:34-53).
    """
    rt, x, dt, want_numpy = _device_record(in_signal)
    return finish(rt, _scaled(rt, x, dt, extraction_type), want_numpy)


def _bandpassed(rt, x, dt, filter_band, sample_rate_hz, filter_order):
    """sosfiltfilt(butter(order, band, fs, "band", "sos"), x) on a device record (reference utilities/picker.py:56-76)."""
    from scipy.signal import butter, sosfilt_zi
    if filter_band[0] < 0 or filter_band[1] > sample_rate_hz / 2:
        raise ValueError(f"Invalid bandpass filter band, {filter_band}, for sample rate {sample_rate_hz}")
    if filter_band[0] >= filter_band[1]:
        raise ValueError(
            f"Invalid bandpass filter band, {filter_band}, the lower bound must be less than the upper bound"
        )
    sos = butter(filter_order, filter_band, fs=sample_rate_hz, btype="band", output="sos")
    ntaps = 2 * sos.shape[0] + 1                                       # scipy.signal.sosfiltfilt's default padding
    ntaps -= min((sos[:, 2] == 0).sum(), (sos[:, 5] == 0).sum())
    out = _driver.filtfilt(rt.reshape(x, (1, x.shape[0])), dt, 3 * int(ntaps), sos=sos, zi=sosfilt_zi(sos), rt=rt)
    return rt.reshape(out, (x.shape[0],))


def apply_bandpass(timeseries: np.ndarray, filter_band: Tuple[float, float], sample_rate_hz: float,
                   filter_order: int = 7) -> np.ndarray:
    """
    Zero-phase Butterworth band-pass in second-order sections (reference utilities/picker.py:56-76).

    :return: filtered signal
    """
    rt, x, dt, want_numpy = _device_record(timeseries)
    return finish(rt, _bandpassed(rt, x, dt, filter_band, sample_rate_hz, filter_order), want_numpy)


def find_peaks_by_extraction_type_with_bandpass(timeseries: np.ndarray, filter_band: Tuple[float, float],
                                                sample_rate_hz: float, filter_order: int = 7,
                                                extraction_type: str = "sigmax", height: Optional[float] = 0.7,
                                                *args) -> np.ndarray:
    """
    Peaks of the band-passed, scaled record that reach ``height`` (reference utilities/picker.py:79-105); the record
    stays on the device from the filter to the peak scan.

    :return: sample positions of the peaks (int64, ascending)
    """
    _no_extra(args)
    rt, x, dt, _ = _device_record(timeseries)
    filtered = _bandpassed(rt, x, dt, filter_band, sample_rate_hz, filter_order)
    return _find_peaks(rt, _scaled(rt, filtered, dt, extraction_type), dt, height)


def _select_by_peak_distance(lib, peaks: np.ndarray, priority: np.ndarray, distance: float) -> np.ndarray:
    """scipy/signal/_peak_finding_utils.pyx::_select_by_peak_distance on the candidate list (host, compiled:
    ``qi_select_peaks_by_distance``): highest priority first, every kept peak removes its neighbours closer than
    ceil(distance) samples.  The visiting order is numpy's own argsort of the priorities, as in scipy."""
    peaks = np.ascontiguousarray(peaks, dtype=np.int64)
    order = np.ascontiguousarray(np.argsort(priority), dtype=np.int64)
    keep = np.empty(peaks.shape[0], dtype=np.uint8)
    rc = lib.qi_select_peaks_by_distance(peaks.ctypes.data, order.ctypes.data, peaks.shape[0], int(np.ceil(distance)),
                                         keep.ctypes.data)
    _lib.check(lib, rc, "qi_select_peaks_by_distance")
    return keep.astype(bool)


def _find_peaks(rt, x, dt, height, distance=None) -> np.ndarray:
    """scipy.signal.find_peaks(x, height=height, distance=distance)[0] with the scan on the device."""
    if distance is not None and distance < 1:
        raise ValueError("`distance` must be greater or equal to 1")
    peaks, priority = _driver.local_maxima(x, dt, height=height, rt=rt)
    if distance is not None and peaks.size:
        peaks = peaks[_select_by_peak_distance(rt.lib, peaks, priority, distance)]
    return peaks


def _no_extra(args):
    if args:   # the reference forwards *args after height=..., which scipy rejects the same way
        raise TypeError("find_peaks() got multiple values for argument 'height'")


def find_peaks_by_extraction_type(timeseries: np.ndarray, extraction_type: str = "sigmax",
                                  height: Optional[float] = 0.7, *args) -> np.ndarray:
    """
    Peaks of the scaled record that reach ``height`` (reference utilities/picker.py:108-120).

    :return: sample positions of the peaks (int64, ascending)
    """
    _no_extra(args)
    rt, x, dt, _ = _device_record(timeseries)
    return _find_peaks(rt, _scaled(rt, x, dt, extraction_type), dt, height)


def find_peaks_with_bits(timeseries: np.ndarray, sample_rate_hz: float, scaling_type: str = "amplitude",
                         threshold_bits: Optional[int] = 1, time_distance_seconds: Optional[float] = 0.1,
                         *args) -> np.ndarray:
    """
    Peaks of log2(|x| + eps) within ``threshold_bits`` of the maximum and at least ``time_distance_seconds`` apart
    (reference utilities/picker.py:123-149).

    :return: sample positions of the peaks (int64, ascending)
    """
    _no_extra(args)
    rt, x, dt, _ = _device_record(timeseries)
    bits = _driver.abs_log2(x, dt, False, eps=get_epsilon(), rt=rt)
    if scaling_type == "log2":
        vmax, _, _, has_nan = _extrema(rt, bits, dt)
        height = (np.nan if has_nan else vmax) - threshold_bits          # np.max propagates NaN
    else:
        vmax, _, _, has_nan = _extrema(rt, x, dt)
        height = (np.nan if has_nan else vmax) - 2 ** threshold_bits
    return _find_peaks(rt, bits, dt, height, distance=int(time_distance_seconds * sample_rate_hz))


def extract_signal_index_with_buffer(sample_rate_hz: float, peak: int, intro_buffer_s: float, outro_buffer_s: float
                                     ) -> Tuple[int, int]:
    """Start and end index of a window around a peak (reference utilities/picker.py:152-166)."""
    if intro_buffer_s < 0 or outro_buffer_s < 0:
        raise ValueError(f"Negative intro_buffer_s or outro_buffer_s, {intro_buffer_s}, {outro_buffer_s}")
    return peak - int(intro_buffer_s * sample_rate_hz), peak + int(outro_buffer_s * sample_rate_hz)


def extract_signal_with_buffer_seconds(timeseries: np.ndarray, sample_rate_hz: float, peak: int,
                                       intro_buffer_s: float, outro_buffer_s: float) -> np.ndarray:
    """The record around a peak, clipped to the record (reference utilities/picker.py:169-192); a view, no copy."""
    intro_index, outro_index = extract_signal_index_with_buffer(sample_rate_hz, peak, intro_buffer_s, outro_buffer_s)
    if intro_index < 0:
        print(f"Warning: intro buffer exceeds the signal length, intro_index: {intro_index}")
        intro_index = 0
    if outro_index > len(timeseries):
        print(f"Warning: outro buffer exceeds the signal length, outro_index: {outro_index}")
        outro_index = len(timeseries)
    return timeseries[intro_index:outro_index]


def find_peaks_to_comb_function(timeseries: np.ndarray, peaks: Union[list, int, np.ndarray]) -> np.ndarray:
    """Comb of ones at the peak positions, zeros elsewhere (reference utilities/picker.py:195-209)."""
    if isinstance(peaks, np.ndarray):
        peaks = peaks.tolist()
    comb_function = np.zeros(len(timeseries))
    comb_function[peaks] = 1
    return comb_function
