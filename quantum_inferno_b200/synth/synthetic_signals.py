"""
``quantum_inferno.synth.synthetic_signals`` on the B200 (reference synth/synthetic_signals.py:53-81, :127-192): the
linear-sweep generators ``chirp_noise_16bit`` / ``chirp_linear_in_noise`` with their white noise, the anti-alias filter
every synthetic generator of the reference ends with, and the Tukey taper.

The reference draws its noise from the unseeded global ``numpy.random`` state (:169), so a noisy record is not
reproducible there.  Here the noise comes from the device generator; the keyword-only ``seed`` makes a record
repeatable (``seed=None`` = fresh entropy, like the reference).  The deterministic part -- sweep, taper, zero padding,
anti-alias filter, float16 rounding -- is parity-tested against the reference with the noise switched off
(``noise_std_loss_bits = inf``: the reference then adds ``normal(0, 0)`` = 0).  Sweep: csrc/qi_synth.cu; filter:
csrc/qi_iir.cu; the taper multiply and the noise add are elementwise operations on the device buffers.
"""
from typing import Optional, Tuple, Union

import numpy as np

from .. import _driver, _iir
from .._runtime import finish, get_runtime


def _std(x):
    """Population standard deviation (np.std) of a device buffer, as a 0-d value of the buffer's own array type."""
    m = x.mean()
    return ((x - m) * (x - m)).mean() ** 0.5


def white_noise_fbits(sig: np.ndarray, std_bit_loss: float, *, seed: Optional[int] = None) -> np.ndarray:
    """
    White noise with zero mean and a standard deviation ``std_bit_loss`` bits below that of the input signal
    (reference synth/synthetic_signals.py:160-169), drawn on the device.  numpy in -> numpy out, CUDA tensor in ->
    tensor out.
    """
    rt = get_runtime()
    want_numpy = not rt.is_device_array(sig)
    x = rt.reshape(rt.asarray(sig, "float64"), (-1,))
    return finish(rt, rt.randn(x.shape, "float64", seed) * (_std(x) / 2.0 ** std_bit_loss), want_numpy)


def _linear_sweep(rt, n_points, sample_rate_hz, frequency_start_hz, frequency_end_hz):
    """``scipy.signal.chirp(t, f0, t[-1], f1, method="linear")`` times the 25 % Tukey taper, float64 on the device:
    cos(2 pi (f0 t + (f1 - f0) / (2 t1) t^2)), t = k / fs, t1 = (n - 1) / fs."""
    t1 = (n_points - 1) / sample_rate_hz
    beta = (frequency_end_hz - frequency_start_hz) / t1
    wf = _driver.synth_chirp(n_points, "float64", 2 * np.pi * frequency_start_hz / sample_rate_hz,
                             half_gamma=np.pi * beta, chirp_scale=sample_rate_hz, rt=rt)[0]
    return wf * rt.asarray(taper_tukey(np.empty(n_points), 0.25), "float64")


def chirp_noise_16bit(duration_points: int = 2 ** 12, sample_rate_hz: float = 80.0, noise_std_loss_bits: float = 4.0,
                      frequency_center_hz: Optional[float] = None, *, seed: Optional[int] = None) -> np.ndarray:
    """
    Chirp with a linear frequency sweep from fc / 2 to fs / 4, tapered, white noise added, anti-alias filtered, float16
    (reference synth/synthetic_signals.py:53-81).
    """
    if not frequency_center_hz:
        frequency_center_hz = 8.0 / (duration_points / sample_rate_hz)
    rt = get_runtime()
    chirp_wf = _linear_sweep(rt, int(duration_points), sample_rate_hz, 0.5 * frequency_center_hz, sample_rate_hz / 4.0)
    noise = rt.randn(chirp_wf.shape, "float64", seed) * (_std(chirp_wf) / 2.0 ** noise_std_loss_bits)
    chirp_white_aa = _antialias_device(rt, rt.reshape(chirp_wf + noise, (1, -1)), "float64", 4)
    return rt.to_numpy(chirp_white_aa)[0].astype(np.float16)


def chirp_linear_in_noise(snr_bits: float, sample_rate_hz: float, duration_s: float, frequency_start_hz: float,
                          frequency_end_hz: float, intro_s: Union[int, float], outro_s: Union[int, float], *,
                          seed: Optional[int] = None, device_out: bool = False) -> Tuple[np.ndarray, np.ndarray]:
    """
    Tapered linear sweep between ``intro_s`` and ``outro_s`` seconds of silence, white noise ``snr_bits`` below the
    standard deviation of the whole record added (reference synth/synthetic_signals.py:127-157).
    ``device_out=True`` leaves the waveform on the device (the input of the transforms of this package).

    :return: waveform, time in seconds
    """
    rt = get_runtime()
    chirp_wf = _linear_sweep(rt, int(sample_rate_hz * duration_s), sample_rate_hz, frequency_start_hz, frequency_end_hz)
    n_in, n_out, n_chirp = int(intro_s * sample_rate_hz), int(outro_s * sample_rate_hz), int(chirp_wf.shape[0])
    sig_wf = rt.zeros((n_in + n_chirp + n_out,), "float64")
    sig_wf[n_in:n_in + n_chirp] = chirp_wf
    synth_wf = sig_wf + rt.randn(sig_wf.shape, "float64", seed) * (_std(sig_wf) / 2.0 ** snr_bits)
    time_s = np.arange(n_in + n_chirp + n_out) / sample_rate_hz
    return (synth_wf if device_out else rt.to_numpy(synth_wf)), time_s


def taper_tukey(sig_or_time: np.ndarray, fraction_cosine: float) -> np.ndarray:
    """Symmetric Tukey window with the size of the input (reference synth/synthetic_signals.py:166-177); a host
    table of window weights, as in the reference."""
    from scipy import signal
    return signal.windows.tukey(M=np.size(sig_or_time), alpha=fraction_cosine, sym=True)


def _antialias_device(rt, x2, dt, filter_order):
    """``filtfilt(*butter(order, 0.5), .)`` over the rows of the device buffer x2 [M, n]."""
    from scipy import signal
    [b, a] = signal.butter(filter_order, 0.5, btype="lowpass")
    sos = _iir.tf2sos_exact(b, a)
    return _driver.filtfilt(x2, dt, 3 * max(len(a), len(b)), sos=sos, zi=_iir.sosfilt_zi(sos), rt=rt)


def antialias_half_nyquist(synth: np.ndarray, filter_order: int = 4) -> np.ndarray:
    """
    Zero-phase Butterworth low-pass with -3 dB at a quarter of the sample rate (reference
    synth/synthetic_signals.py:180-192): ``filtfilt(*butter(order, 0.5), synth)`` with the two recursions run on the
    device (csrc/qi_iir.cu).  numpy in -> numpy out, CUDA tensor in -> tensor out; 2-D [channels, points] accepted.
    """
    rt = get_runtime()
    want_numpy = not rt.is_device_array(synth)
    name = str(synth.dtype).replace("torch.", "") if not want_numpy else np.asarray(synth).dtype.name
    dt = name if name in ("float32", "float64") else "float64"
    x = rt.asarray(synth, dt)
    lead = tuple(int(v) for v in x.shape[:-1])
    x2 = rt.reshape(x, (int(np.prod(lead)) if lead else 1, int(x.shape[-1])))
    out = _antialias_device(rt, x2, dt, filter_order)
    return finish(rt, rt.reshape(out, lead + (int(x.shape[-1]),)), want_numpy)
