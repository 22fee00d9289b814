"""
Host-side planning (pure numpy, float64): turns the reference's call arguments into the small
per-band parameter tables the kernels consume.  No GPU, no torch -- so it is testable anywhere.
"""
import numpy as np

from . import scales_dyadic as scales
from ._lib import ATOM_BAND

# N/s below which the Gaussian has NOT decayed at the record edge and the truncated atom's exact
# spectrum is used instead of the closed form.  erfc(x/sqrt(2)) at x = N/(2s).
ANALYTIC_MIN_POINTS_PER_SCALE = {"float64": 15.0, "float32": 10.0}


def wavelet_amplitude(scale_atom):
    """(amp_canonical, amp_unit_spectrum) -- reference styx_cwt.py:29-40, expression kept as is."""
    amp_canonical = (np.pi * scale_atom ** 2) ** (-1 / 4)
    amp_unit_spectrum = (4 * np.pi * scale_atom ** 2) ** (-1 / 4) * amp_canonical
    return amp_canonical, amp_unit_spectrum


def dictionary_amplitude(scale_atom, dictionary_type):
    """Amplitude selection of styx_cwt.py:137-142: 'spect', 'unit', anything else -> canonical."""
    amp_canonical, amp_unit_spectrum = wavelet_amplitude(scale_atom)
    if dictionary_type == "spect":
        return amp_unit_spectrum
    if dictionary_type == "unit":
        return 1.0 if np.isscalar(scale_atom) else np.ones(np.shape(scale_atom))
    return amp_canonical


def gabor_bands(band_order_nth, n_points, frequency_hz, frequency_sample_rate_hz, dictionary_type="norm",
                dtype_name="float64", spectrum="auto"):
    """Band table for the styx_cwt Gabor dictionary.

    spectrum: 'auto' (closed form where exact, table for truncated low bands), 'table' (always transform the
    time-domain atom, the literal restatement of scipy fftconvolve) or 'analytic'.
    Returns (bands[ATOM_BAND], scale, omega, amp)."""
    f = np.atleast_1d(np.asarray(frequency_hz, dtype=np.float64))
    scale, omega = scales.scale_from_frequency_hz(band_order_nth, f, frequency_sample_rate_hz)
    amp = np.broadcast_to(np.asarray(dictionary_amplitude(scale, dictionary_type), dtype=np.float64), f.shape)
    bands = np.zeros(len(f), dtype=ATOM_BAND)
    bands["omega"] = omega
    bands["p_re"] = 0.5 / scale ** 2
    bands["amp"] = amp
    if spectrum == "table":
        bands["analytic"] = 0
    elif spectrum == "analytic":
        bands["analytic"] = 1
    else:
        bands["analytic"] = (n_points / scale >= ANALYTIC_MIN_POINTS_PER_SCALE[dtype_name]).astype(np.int32)
    return bands, scale, omega, amp
