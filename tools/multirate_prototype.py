"""
Numpy model of the multirate (fp32-accuracy) CWT algorithm, used to validate the design before/alongside the
CUDA kernels (csrc/qi_mr_*.cu): half-band decimation pyramid of the record, each band evaluated at the deepest
level whose alias-free band [0, pi/2] still holds its whole Gaussian response, then half-band interpolation back
to the full rate.  Run:  python tools/multirate_prototype.py
"""
import os
import re
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from quantum_inferno_b200 import _plan, scales_dyadic as sc  # noqa: E402

KAPPA = 4.8          # half-width of the band response kept, in units of 1/s
HALO = 16


def load_taps():
    txt = open(os.path.join(ROOT, "quantum_inferno_b200", "csrc", "qi_halfband_coeffs.h")).read()
    ntaps = [int(v) for v in re.search(r"qi_hb_ntaps\[.*?\] = \{(.*?)\}", txt).group(1).split(",")]
    rows = re.findall(r"^\s*\{(.*?)\},", txt, flags=re.M)
    return [np.array([float(v) for v in r.split(",")])[:n] for r, n in zip(rows, ntaps)]


TAPS = load_taps()


def taps_for_class(j):
    return TAPS[min(j, len(TAPS)) - 1]


def decimate(x_lo_halo, n_next):
    """x on q in [-HALO, n+HALO) -> next level on [-HALO, n_next+HALO); zero outside the stored range."""
    c = TAPS[0]
    n_in = len(x_lo_halo)
    pad = 2 * HALO + 2 * len(c) + 4
    xp = np.concatenate([np.zeros(pad), x_lo_halo, np.zeros(pad)])
    q = np.arange(-HALO, n_next + HALO)
    centre = 2 * q + HALO + pad                      # index of sample 2q in xp
    out = 0.5 * xp[centre]
    for i, ci in enumerate(c, start=1):
        out += ci * (xp[centre + (2 * i - 1)] + xp[centre - (2 * i - 1)])
    return out


def interpolate(y, cls):
    """y on [-h, n+h) at level m+1 -> level m on [-(h-n_t)*2.., ...]; returns (array, new_halo)."""
    c = taps_for_class(cls)
    nt = len(c)
    h_in = interpolate.halo
    lo, hi = nt - 1, len(y) - nt                    # q range with full support for the odd sample 2q+1
    q = np.arange(lo, hi)
    even = y[q]
    odd = np.zeros(len(q), dtype=complex)
    for i, ci in enumerate(c, start=1):
        odd += 2.0 * ci * (y[q + i] + y[q - i + 1])      # interpolator = 2 x the half-band low-pass
    out = np.empty(2 * len(q), dtype=complex)
    out[0::2] = even
    out[1::2] = odd
    interpolate.halo = 2 * (h_in - lo)
    return out


def band_levels(omega, scale, n_points):
    m = omega * scale
    cap = max(0, int(np.log2(n_points)) - 10)
    lv = np.floor(np.log2((np.pi / 2) / (omega * (1 + KAPPA / m)))).astype(int)
    return np.clip(lv, 0, cap), cap


def multirate_cwt(x, order, fs, dictionary="norm"):
    n = len(x)
    f = sc.log_frequency_hz_from_fft_points(fs, n, order)
    bands, scale, omega, amp = _plan.gabor_bands(order, n, f, fs, dictionary)
    levels, cap = band_levels(omega, scale, n)
    # pyramid
    pyr = [np.concatenate([np.zeros(HALO), x, np.zeros(HALO)])]
    for lvl in range(1, cap + 1):
        pyr.append(decimate(pyr[-1], n >> lvl))
    out = np.empty((len(f), n), dtype=complex)
    for b in range(len(f)):
        lvl = levels[b]
        xl = pyr[lvl]
        n_l = n >> lvl
        step = 2 ** lvl
        # kernel at the level rate, lags d in [-(n_l + 2 HALO), +..]; truncated like the reference's atom
        dmax = n_l + 2 * HALO
        d = np.arange(-dmax, dmax + 1)
        t = step * d - 0.5
        ker = step * amp[b] * np.exp(-0.5 * (t / scale[b]) ** 2) * np.exp(1j * omega[b] * t)
        ker[np.abs(t) > (n - 1) / 2] = 0.0
        full = np.convolve(xl, ker) if len(xl) < 4096 else fftconv(xl, ker)
        # output index q (level rate) <-> full index (q + HALO) + dmax
        q = np.arange(-HALO, n_l + HALO)
        y = full[q + HALO + dmax]
        interpolate.halo = HALO
        for m in range(lvl - 1, -1, -1):
            y = interpolate(y, lvl - m)
        h = interpolate.halo
        out[b] = y[h:h + n]
    return f, levels, out


def fftconv(a, b):
    L = 1 << int(np.ceil(np.log2(len(a) + len(b) - 1)))
    return np.fft.ifft(np.fft.fft(a, L) * np.fft.fft(b, L))[:len(a) + len(b) - 1]


if __name__ == "__main__":
    from oracle import qi_oracle as orc
    fs = 800.0
    for order, logn in [(3, 12), (3, 14), (3, 16), (6, 14), (12, 14), (1, 13)]:
        n = 1 << logn
        k = np.arange(n)
        x = (np.cos(2 * np.pi * 60.0 / fs * k) + 0.5 * np.cos(2 * np.pi * (k / fs + 0.5 * (199.0 / (n / fs)) * (k / fs) ** 2))
             + np.random.default_rng(3).standard_normal(n) / 16)
        fr, tr, cr = orc.cwt_complex_any_scale_pow2(order, x, fs)
        f, lv, c = multirate_cwt(x, order, fs)
        pr, p = np.abs(cr) ** 2, np.abs(c) ** 2
        band_err = np.max(np.abs(c - cr), axis=1) / np.max(np.abs(cr))
        print(f"order {order} n=2^{logn}: B={len(f)} levels {lv.min()}..{lv.max()} "
              f"power rel L2 {np.linalg.norm(p - pr) / np.linalg.norm(pr):.2e} "
              f"max amp err {band_err.max():.2e} (band {band_err.argmax()}, level {lv[band_err.argmax()]}) "
              f"per-band-rel max {np.max(np.max(np.abs(c - cr), axis=1) / np.max(np.abs(cr), axis=1)):.2e}")
