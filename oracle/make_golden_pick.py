"""
make_golden_pick.py -- generates tests/golden/pick.npz by running the UNMODIFIED reference
(quantum_inferno.utilities.sampling / .picker imported from /root/reference, with the installed scipy's
find_peaks behind it) on seeded inputs.  SURVEY 8(f) rank 3.  Build container only; test infrastructure.
    python oracle/make_golden_pick.py
"""
import os
import sys

import numpy as np

REF = os.environ.get("QI_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
from quantum_inferno.utilities import picker, sampling  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
FACTORS = (2, 3, 7, 16, 32, 100, 128, 129, 300, 1000)
METHODS = ("average", "median", "max", "min", "nth")


def burst_record(n=6000, fs=800.0, seed=5):
    """Five Gaussian-windowed tone bursts of different heights on a noise floor, plus two exact flat tops."""
    rng = np.random.default_rng(seed)
    k = np.arange(n)
    x = 0.02 * rng.standard_normal(n)
    for c, a, w in [(600, 1.0, 40.0), (1500, 0.8, 25.0), (1560, 0.75, 25.0), (3000, 0.55, 60.0), (4800, 0.95, 30.0)]:
        x += a * np.exp(-0.5 * ((k - c) / w) ** 2) * np.cos(2 * np.pi * 60.0 / fs * (k - c))
    x[2000:2004] = 0.9          # plateau of four equal samples (even length: midpoint rounds down)
    x[2500:2503] = 0.85         # plateau of three
    return x


def main():
    rng = np.random.default_rng(42)
    d = {}
    plane = rng.standard_normal((5, 3001)) ** 2
    plane[1, 17] = np.nan
    plane[3, 2990:] = 2.5                                                  # ties
    d["plane"] = plane
    d["factors"] = np.array(FACTORS)
    for dt in ("float64", "float32"):
        p = plane.astype(dt)
        for f in FACTORS:
            for m in METHODS:
                d[f"sub2d_{dt}_{f}_{m}"] = sampling.subsample_2d(p, f, m)
    row = rng.standard_normal(4099)
    d["row"] = row
    for f in (2, 5, 64, 200):
        for m in METHODS:
            y, rate = sampling.subsample(row, 800.0, f, m)
            d[f"sub1d_{f}_{m}"] = y
            d[f"sub1d_{f}_{m}_rate"] = np.array(rate)

    x = burst_record()
    d["x"] = x
    for et in picker.EXTRACTION_TYPE:
        d[f"scaled_{et}"] = picker.scale_signal_by_extraction_type(x, et)
        for h in (0.7, 0.3):
            d[f"peaks_{et}_{h}"] = picker.find_peaks_by_extraction_type(x, et, h)
    for st in picker.INPUT_SCALE_TYPE:
        for tb in (1, 3):
            for dist in (0.1, 0.01, 0.5):
                d[f"bits_{st}_{tb}_{dist}"] = picker.find_peaks_with_bits(x, 800.0, st, tb, dist)
    noise = rng.standard_normal(20000)
    d["noise"] = noise
    d["noise_peaks_bits"] = picker.find_peaks_with_bits(noise, 800.0, "log2", 2, 0.05)
    d["noise_peaks_sigmax"] = picker.find_peaks_by_extraction_type(noise, "sigmax", 0.5)
    np.savez_compressed(os.path.join(OUT, "pick.npz"), **d)
    print("pick.npz", os.path.getsize(os.path.join(OUT, "pick.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
