"""
Block sub-sampling of records and time-frequency planes on the B200 -- drop-in for the reductions of
``quantum_inferno.utilities.sampling`` (reference utilities/sampling.py:13-50 ``subsample`` and :87-120
``subsample_2d``; same names, arguments, printed warnings and return values).

This is the step AFTER the time-frequency path: a [bands, time] power plane that lives in HBM is reduced to a
displayable mesh (``scales_dyadic.DEFAULT_MESH_POW2_PIXELS``) without leaving the device.  All five methods run in
csrc/qi_pick.cu (one streaming read of the plane); numpy in -> numpy out, CUDA tensor in -> tensor out.  The mean
accumulates in fp64 (above the four float32 samples of one 128-bit load); "median" / "max" / "min" / "nth" are bit-exact.  float32 and float64 are computed as they are,
any other dtype is converted to float64 first.  ``decimate_timeseries`` / ``decimate_timeseries_collection``
(:123-146, ``scipy.signal.decimate(zero_phase=True)``) run scipy's own order-8 Chebyshev cascade through the blocked
IIR scan of csrc/qi_iir.cu and keep every q-th sample; the interpolating resamplers (:53-84) are not provided.
"""
from typing import Tuple

import numpy as np

from .. import _driver
from .._runtime import finish, get_runtime

SUBSAMPLE_METHODS = ["average", "median", "max", "min", "nth"]


def _dtype_of(rt, x):
    if rt.is_device_array(x):
        name = str(x.dtype).replace("torch.", "")
    else:
        name = np.asarray(x).dtype.name
    return name if name in ("float32", "float64") else "float64"


def _checked_method(method: str) -> str:
    if method not in SUBSAMPLE_METHODS:
        print(f"Warning: method {method} not recognized, using 'nth' method")
        return "nth"
    return method


def subsample(timeseries: np.ndarray, sample_rate_hz: float, subsample_factor: int, method: str = "nth"
              ) -> Tuple[np.ndarray, float]:
    """
    Subsample a time series (reference utilities/sampling.py:13-50): every ``subsample_factor`` samples are replaced
    by their mean / median / max / min, or every nth sample is kept.  The tail that does not fill a group is dropped
    (kept for "nth").  A factor below 2 returns the input unchanged with the reference's warning.

    :return: subsampled signal and new sample rate
    """
    if subsample_factor < 2:
        print(f"Warning: subsample factor is less than 2, returning the original signal")
        return timeseries, sample_rate_hz
    new_sample_rate = sample_rate_hz / subsample_factor
    method = _checked_method(method)
    rt = get_runtime(timeseries)
    dt = _dtype_of(rt, timeseries)
    want_numpy = not rt.is_device_array(timeseries)
    x = rt.asarray(timeseries, dt)
    if x.ndim != 1:
        raise ValueError("timeseries must be 1-D")
    out = _driver.subsample(rt.reshape(x, (1, x.shape[0])), subsample_factor, method, dt, rt=rt)
    return finish(rt, rt.reshape(out, (out.shape[1],)), want_numpy), new_sample_rate


def subsample_2d(array: np.ndarray, subsample_factor: int, method: str = "nth") -> np.ndarray:
    """
    Subsample a 2-D array along its second axis (reference utilities/sampling.py:87-120).  A leading batch axis
    ([channels, bands, time]) is accepted as an extension.

    :return: subsampled array, [rows, floor(n / factor)] (ceil for "nth")
    """
    if subsample_factor < 2:
        print(f"Warning: subsample factor is less than 2, returning the original signal")
        return array
    method = _checked_method(method)
    rt = get_runtime(array)
    dt = _dtype_of(rt, array)
    want_numpy = not rt.is_device_array(array)
    x = rt.asarray(array, dt)
    if x.ndim not in (2, 3):
        raise ValueError("array must be 2-D [rows, points] (or 3-D [channels, rows, points])")
    lead = tuple(int(s) for s in x.shape[:-1])
    out = _driver.subsample(rt.reshape(x, (int(np.prod(lead)), x.shape[-1])), subsample_factor, method, dt, rt=rt)
    return finish(rt, rt.reshape(out, lead + (out.shape[1],)), want_numpy)


def _decimate_device(rt, x2, dt, q: int):
    """scipy.signal.decimate(x, q, ftype="iir", zero_phase=True) along the last axis of a device buffer [M, n]:
    sosfiltfilt with cheby1(8, 0.05, 0.8 / q) then every q-th sample (scipy/signal/_signaltools.py::decimate)."""
    import operator

    from scipy.signal import cheby1, sosfilt_zi
    q = operator.index(q)
    sos = cheby1(8, 0.05, 0.8 / q, output="sos")
    ntaps = 2 * sos.shape[0] + 1
    ntaps -= min((sos[:, 2] == 0).sum(), (sos[:, 5] == 0).sum())
    y = _driver.filtfilt(x2, dt, 3 * int(ntaps), sos=sos, zi=sosfilt_zi(sos), rt=rt)
    return _driver.subsample(y, q, "nth", dt, rt=rt) if q > 1 else y


def decimate_timeseries(timeseries: np.ndarray, decimation_factor: int) -> np.ndarray:
    """
    Anti-aliased decimation of a time series (reference utilities/sampling.py:123-133; 28 samples or longer).

    :return: decimated signal
    """
    rt = get_runtime(timeseries)
    dt = _dtype_of(rt, timeseries)
    want_numpy = not rt.is_device_array(timeseries)
    x = rt.asarray(timeseries, dt)
    if x.ndim != 1:
        raise ValueError("timeseries must be 1-D")
    out = _decimate_device(rt, rt.reshape(x, (1, x.shape[0])), dt, decimation_factor)
    return finish(rt, rt.reshape(out, (out.shape[1],)), want_numpy)


def decimate_timeseries_collection(timeseries_collection: np.ndarray, decimation_factor: int) -> np.ndarray:
    """
    Anti-aliased decimation of several time series of equal length at once (reference utilities/sampling.py:137-146).

    :return: decimated signals, [rows, ceil(n / q)]
    """
    rt = get_runtime(timeseries_collection)
    dt = _dtype_of(rt, timeseries_collection)
    want_numpy = not rt.is_device_array(timeseries_collection)
    x = rt.asarray(timeseries_collection, dt)
    if x.ndim != 2:
        raise ValueError("timeseries_collection must be 2-D [rows, points]")
    return finish(rt, _decimate_device(rt, x, dt, decimation_factor), want_numpy)
