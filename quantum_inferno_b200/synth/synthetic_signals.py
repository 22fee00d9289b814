"""
The deterministic conditioning steps of ``quantum_inferno.synth.synthetic_signals`` on the B200 (reference
synth/synthetic_signals.py:166-192): the anti-alias filter every synthetic generator of the reference ends with, and
the Tukey taper.  The generators themselves draw unseeded ``numpy.random`` noise (:169) and are not reproduced.
"""
import numpy as np

from .. import _driver, _iir
from .._runtime import finish, get_runtime


def taper_tukey(sig_or_time: np.ndarray, fraction_cosine: float) -> np.ndarray:
    """Symmetric Tukey window with the size of the input (reference synth/synthetic_signals.py:166-177); a host
    table of window weights, as in the reference."""
    from scipy import signal
    return signal.windows.tukey(M=np.size(sig_or_time), alpha=fraction_cosine, sym=True)


def antialias_half_nyquist(synth: np.ndarray, filter_order: int = 4) -> np.ndarray:
    """
    Zero-phase Butterworth low-pass with -3 dB at a quarter of the sample rate (reference
    synth/synthetic_signals.py:180-192): ``filtfilt(*butter(order, 0.5), synth)`` with the two recursions run on the
    device (csrc/qi_iir.cu).  numpy in -> numpy out, CUDA tensor in -> tensor out; 2-D [channels, points] accepted.
    """
    from scipy import signal
    [b, a] = signal.butter(filter_order, 0.5, btype="lowpass")
    rt = get_runtime()
    want_numpy = not rt.is_device_array(synth)
    name = str(synth.dtype).replace("torch.", "") if not want_numpy else np.asarray(synth).dtype.name
    dt = name if name in ("float32", "float64") else "float64"
    x = rt.asarray(synth, dt)
    lead = tuple(int(v) for v in x.shape[:-1])
    x2 = rt.reshape(x, (int(np.prod(lead)) if lead else 1, int(x.shape[-1])))
    sos = _iir.tf2sos_exact(b, a)
    out = _driver.filtfilt(x2, dt, 3 * max(len(a), len(b)), sos=sos, zi=_iir.sosfilt_zi(sos), rt=rt)
    return finish(rt, rt.reshape(out, lead + (int(x.shape[-1]),)), want_numpy)
