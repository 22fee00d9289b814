"""
Host-side planning (pure numpy, float64): turns the reference's call arguments into the small
per-band parameter tables the kernels consume.  No GPU, no torch -- so it is testable anywhere.
"""
import os

import numpy as np

from . import scales_dyadic as scales
from ._lib import ATOM_BAND

# N/s below which the Gaussian has NOT decayed at the record edge and the truncated atom's exact
# spectrum is used instead of the closed form.  erfc(x/sqrt(2)) at x = N/(2s).
ANALYTIC_MIN_POINTS_PER_SCALE = {"float64": 15.0, "float32": 10.0}
ANALYTIC_MIN_SCALE = 1.0


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
        # closed form = Gaussian + 5 Poisson alias terms: exact only if the atom has decayed at the record edge
        # and is at least one sample wide (narrower atoms -- orders <= 1 near Nyquist -- need more alias terms)
        ok = (n_points / scale >= ANALYTIC_MIN_POINTS_PER_SCALE[dtype_name]) & (scale >= ANALYTIC_MIN_SCALE)
        bands["analytic"] = ok.astype(np.int32)
    return bands, scale, omega, amp


def nearest_fft_bin(frequency, n_points, sample_interval):
    """``np.abs(np.fft.fftfreq(n_points, sample_interval) - frequency).argmin()`` -- the Stockwell shift index
    of reference styx_stx.py:167,233 -- evaluated on a handful of candidate bins with the same float64
    expressions (bin value = signed_bin * (1/(n*d)), first minimum wins) instead of on the whole bin table,
    so it stays O(1) for 2^28-point records."""
    n = int(n_points)
    val = 1.0 / (n * sample_interval)
    n_pos = (n - 1) // 2 + 1                      # indices [0, n_pos) hold bins 0..n_pos-1, the rest are negative

    def bin_value(idx):
        return (idx if idx < n_pos else idx - n) * val

    cand = {0, n_pos - 1, min(n_pos, n - 1), n - 1}
    guess = frequency / val
    if np.isfinite(guess):
        for base in (int(np.floor(guess)), int(np.ceil(guess))):
            for signed in (base - 1, base, base + 1):
                if 0 <= signed < n_pos:
                    cand.add(signed)
                elif -(n // 2) <= signed < 0:
                    cand.add(signed + n)
    return min(sorted(cand), key=lambda idx: (abs(bin_value(idx) - frequency), idx))


def stx_bands(band_order_nth, n_points, frequency_sample_rate_hz):
    """Band table of styx_stx.stx_complex_any_scale_pow2 (reference styx_stx.py:206-233).
    sigma uses the exact band centre, the shift index is the first argmin over the fft bins.
    Returns (frequency_stx_hz, bands[STX_BAND])."""
    from ._lib import STX_BAND
    f_stx = scales.log_frequency_hz_from_fft_points(frequency_sample_hz=frequency_sample_rate_hz,
                                                    fft_points=n_points, scale_order=band_order_nth)
    omega_stx = 2 * np.pi * f_stx / frequency_sample_rate_hz
    bands = np.zeros(len(f_stx), dtype=STX_BAND)
    bands["sigma"] = scales.cycles_from_order(scale_order=band_order_nth) / omega_stx
    bands["shift"] = [nearest_fft_bin(f, n_points, 1 / frequency_sample_rate_hz) for f in f_stx]
    return f_stx, bands


# ----------------------------------------------------------------------------- STFT
def periodic_window(kind, param, n_points):
    """The DFT-even window scipy.signal.get_window((kind, param), n_points) returns: the symmetric
    (n_points+1)-point window without its last sample.  kind: 'tukey' (param = alpha) or 'gaussian'
    (param = sigma in samples) -- the two windows the reference asks for (styx_fft.py:178,218,257)."""
    m = n_points + 1
    k = np.arange(m, dtype=np.float64)
    if kind == "tukey":
        alpha = float(param)
        if alpha <= 0:
            w = np.ones(m)
        elif alpha >= 1.0:
            w = 0.5 - 0.5 * np.cos(2.0 * np.pi * k / (m - 1))
        else:
            edge = int(np.floor(alpha * (m - 1) / 2.0))
            rise = 0.5 * (1 + np.cos(np.pi * (-1 + 2.0 * k[:edge + 1] / alpha / (m - 1))))
            fall = 0.5 * (1 + np.cos(np.pi * (-2.0 / alpha + 1 + 2.0 * k[m - edge - 1:] / alpha / (m - 1))))
            w = np.concatenate((rise, np.ones(m - 2 * edge - 2), fall))
    elif kind == "gaussian":
        w = np.exp(-0.5 * ((k - (m - 1.0) / 2.0) / float(param)) ** 2)
    else:
        raise ValueError(f"unsupported window {kind!r}")
    return w[:-1]


def stft_frames(n_points, nperseg, noverlap, boundary_zeros=True, padded=True):
    """Frame bookkeeping of scipy's _spectral_helper: returns (n_frames, pad_left, extended_length)."""
    hop = nperseg - noverlap
    if hop <= 0:
        raise ValueError("noverlap must be less than nperseg.")
    pad_left = nperseg // 2 if boundary_zeros else 0
    ext = n_points + 2 * pad_left
    if padded:
        ext += (-(ext - nperseg) % hop) % nperseg
    n_frames = (ext - noverlap) // hop
    return n_frames, pad_left, ext


def stft_time_axis(ext_length, nperseg, noverlap, fs, boundary_zeros=True):
    t = np.arange(nperseg / 2, ext_length - nperseg / 2 + 1, nperseg - noverlap) / float(fs)
    if boundary_zeros:
        t -= (nperseg / 2) / fs
    return t


# ----------------------------------------------------------------------------- multirate fp32 CWT
# Half-width (in 1/s) of the band response that must sit inside a level's alias-free band [0, pi/2] when the band's level is
# chosen.  What has to hold is that the LAST decimation stage of the band's level neither droops nor aliases where the band
# still answers: max over theta of |H(theta) - 1| * exp(-((theta - omega) s)^2 / 2) for the 7-tap minimax half-band filter.
# That figure is the filter's own pass-band ripple (9.4e-7) for every half-width from 4.8 down to 2.0 (orders 1.5 ... 24;
# the filter is still within 4e-5 of 1 at 1.06 pi/4) and rises below 2.0; the envelope decimation has its own check in
# mr_plan (csrc/qi_mr.cu), unaffected down to 2.0.  2.4 instead of the 4.8 of round 1 moves one band per order from the
# full-rate level 0 -- the expensive one: convolution at the full rate + a separate information pass -- to level 1.
MR_KAPPA = float(os.environ.get("QI_MR_KAPPA", "2.4"))     # the environment override is for measurements only
MR_PASS = np.pi / 2      # alias-free band of every pyramid level, in radians at that level's rate
MR_MIN_LOG2_POINTS = 13


def multirate_supported(n_points, scale):
    """The multirate path needs a 2^m record (m >= 13) and atoms at least one sample wide."""
    n = int(n_points)
    return n >= (1 << MR_MIN_LOG2_POINTS) and (n & (n - 1)) == 0 and bool(np.all(np.asarray(scale) >= ANALYTIC_MIN_SCALE))


def multirate_bands(band_order_nth, n_points, frequency_hz, frequency_sample_rate_hz, dictionary_type="norm"):
    """Level assignment of the multirate CWT: band b runs at level l_b = the largest l with
    (omega_b + kappa/s_b) * 2^l <= pi/2, capped so the deepest level still has 1024 samples.
    Returns (bands[MR_BAND] ascending in frequency, scale, omega, amp)."""
    from ._lib import MR_BAND
    f = np.atleast_1d(np.asarray(frequency_hz, dtype=np.float64))
    scale, omega = scales.scale_from_frequency_hz(band_order_nth, f, frequency_sample_rate_hz)
    amp = np.broadcast_to(np.asarray(dictionary_amplitude(scale, dictionary_type), dtype=np.float64), f.shape)
    cap = int(np.log2(n_points)) - 10
    upper_edge = omega * (1.0 + MR_KAPPA / (omega * scale))
    level = np.floor(np.log2(MR_PASS / upper_edge)).astype(np.int64)
    bands = np.zeros(len(f), dtype=MR_BAND)
    bands["omega"], bands["scale"], bands["amp"] = omega, scale, amp
    bands["level"] = np.clip(level, 0, cap)
    return bands, scale, omega, amp
