"""
numpy model of a multirate (fp32-tolerance) Stockwell transform -- design evidence for the next step listed in
DESIGN.md section 5 ("Stockwell on the multirate machinery").  Not used by the package.

Reference (quantum_inferno/styx_stx.py:213-234):  tfr[b] = ifft( X[(k + shift_b) mod n] * exp(-0.5 sigma_b^2 w_k^2) ).
The product is exactly zero beyond |k| > kmax_b = u0 / (sigma_b 2 pi / n) in float32 (u0 = 14.5), so the voice is a
BASEBAND signal of two-sided bandwidth 2 kmax_b / n.  Taking the K_b = 2^m >= 4 kmax_b bins around zero and an
inverse transform of length K_b gives its samples at the decimated rate n / K_b exactly (2x oversampled); zero-phase
half-band interpolation (the minimax stages of tools/design_halfband.py, circular here) brings them back to the full
rate.  The script reports, per band, the decimation, the error against the full-length inverse transform and the
work relative to it.

    python tools/stx_multirate_prototype.py [log2n] [order]
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from quantum_inferno_b200 import _plan  # noqa: E402
from tools.design_halfband import design  # noqa: E402

U_ZERO = 14.5          # exp(-0.5 u^2) == 0 in float32 beyond this


def halfband_taps(cls, cache={}):
    """Shortest minimax half-band with 1e-6 pass-band error for a signal inside |theta| <= pi / 2^cls at the low rate."""
    if cls not in cache:
        for n in range(1, 40):
            c, err = design(np.pi / 2 ** (cls + 1), n)
            if err <= 1e-6:
                cache[cls] = c
                break
    return cache[cls]


def interpolate2_circular(y, cls):
    """Zero-stuff by two and apply the zero-phase half-band (gain 2), circular boundary."""
    c = halfband_taps(cls)
    up = np.zeros(2 * len(y), dtype=y.dtype)
    up[::2] = y                                       # even outputs: the centre tap (1/2 * 2)
    odd = np.zeros(len(y), dtype=y.dtype)
    for i, ci in enumerate(c):                        # odd outputs: 2 * sum_i c_i (y[q - i] + y[q + 1 + i])
        odd += 2.0 * ci * (np.roll(y, i) + np.roll(y, -(i + 1)))
    up[1::2] = odd
    return up


def main():
    logn = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    order = float(sys.argv[2]) if len(sys.argv) > 2 else 3.0
    n, fs = 1 << logn, 800.0
    k = np.arange(n)
    x = np.cos(2 * np.pi * 60.0 / fs * k) + 0.5 * np.cos(2 * np.pi * (1.0 * k / fs + 0.5 * (199.0 / (n / fs)) * (k / fs) ** 2)) \
        + np.random.default_rng(1).standard_normal(n) / 16.0
    freq, bands = _plan.stx_bands(order, n, fs)
    sigma, shift = bands["sigma"], bands["shift"]
    X = np.fft.fft(x)
    ks = np.fft.fftfreq(n, 1.0 / n)                   # signed bins
    tot_full = tot_fast = 0.0
    worst = 0.0
    print(f"n = 2^{logn}, order {order:g}, {len(freq)} bands")
    for b in range(len(freq)):
        q = sigma[b] * 2 * np.pi / n
        win = np.exp(-0.5 * (q * ks) ** 2)
        full = np.fft.ifft(np.roll(X, -int(shift[b])) * win)
        kmax = int(np.ceil(U_ZERO / q)) + 2
        K = 64
        while K < 4 * kmax:
            K *= 2
        if K >= n:
            fast, D = full, 1
        else:
            sel = np.r_[0:K // 2, n - K // 2:n]       # bins [-K/2, K/2) in fft order
            dec = np.fft.ifft((np.roll(X, -int(shift[b])) * win)[sel]) * (K / n)
            D = n // K
            fast, stage = dec, 1
            while len(fast) < n:
                fast = interpolate2_circular(fast, stage)         # the signal sits inside pi / 2^stage at each low rate
                stage += 1
        err = np.linalg.norm(fast - full) / np.linalg.norm(full)
        worst = max(worst, err)
        # HBM traffic per output cell (complex64): the full-length path reads and writes every point in each of its two
        # passes (32 B); the decimated path writes the cell once and moves 16 B per decimated sample
        tot_full += 32.0
        tot_fast += 32.0 if D == 1 else 8.0 + 16.0 / D
        if b % max(1, len(freq) // 12) == 0 or b == len(freq) - 1:
            print(f"  band {b:3d}  f = {freq[b]:9.4f} Hz  kmax = {kmax:7d}  decimation {D:6d}  rel L2 error {err:.2e}")
    print(f"worst relative L2 error {worst:.2e} (north-star float32 tolerance 1e-4)")
    print(f"HBM traffic model: {tot_fast / len(freq):.1f} B per cell against {tot_full / len(freq):.1f} B for the full-length passes "
          f"(cut-off u0 = {U_ZERO}: the float32 underflow point; u0 = 5.2 already meets 1e-4 and decimates "
          f"{int(np.sum(4 * (np.ceil(5.2 / (sigma * 2 * np.pi / n)) + 2) < n))} of {len(freq)} bands)")


if __name__ == "__main__":
    main()
