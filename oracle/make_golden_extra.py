"""
make_golden_extra.py -- generates tests/golden/extra.npz by running the UNMODIFIED reference (imported from
/root/reference): the thinly covered rows of SURVEY 8(a) at more sizes -- a5 (atoms of the three dictionaries), a9
(general Stockwell transform at a multi-pass size), a14 (every attribute of ShannonTDR / ShannonFFT, record lengths that
are and are not powers of two).  Build container only; test infrastructure.
    python oracle/make_golden_extra.py
"""
import os
import sys

import numpy as np

REF = os.environ.get("QI_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
from quantum_inferno import styx_cwt, styx_stx, tfr_info  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
FS = 800.0
ATOM_CASES = [(3, 512, "norm"), (3, 1000, "spect"), (6, 2048, "unit"), (12, 1000, "norm"), (6, 512, "spect"), (3, 2048, "unit")]
ATOM_FREQS = np.array([2.5, 40.0, 310.0])
INFO_LENGTHS = [1000, 777, 1 << 13]


def record(n, seed):
    k = np.arange(n)
    return np.cos(2 * np.pi * 60.0 / FS * k) + 0.25 * np.random.default_rng(seed).standard_normal(n)


def main():
    d = {}
    for i, (order, n, dic) in enumerate(ATOM_CASES):
        atoms, t_s, scale, omega, amp = styx_cwt.wavelet_centered_4cwt(order, n, ATOM_FREQS, FS, dic)
        d[f"atoms{i}"], d[f"atoms{i}_t"] = atoms, t_s
        d[f"atoms{i}_scale"], d[f"atoms{i}_omega"], d[f"atoms{i}_amp"] = scale[:, 0], omega[:, 0], amp[:, 0]
    # a9 at 2048 samples (two FFT passes), a band subset so the fixture stays small: linear and geometric grids
    x = record(2048, 5)
    d["stx_x"] = x
    for tag, kw in [("lin", dict(frequency_min=10.0, frequency_max=100.0, frequency_step=5.0)),
                    ("geo", dict(frequency_min=10.0, frequency_max=100.0, is_geometric=True, scale_order_input=3.0))]:
        tfr, psd, f, f_fft, win = styx_stx.tfr_stx_fft(x, 1 / FS, n_fft_in=2048, **kw)
        d[f"stx_{tag}_tfr"], d[f"stx_{tag}_psd"], d[f"stx_{tag}_f"] = tfr, psd, f
        d[f"stx_{tag}_ffft"], d[f"stx_{tag}_win0"] = f_fft, win[::4]
    for n in INFO_LENGTHS:
        xr = record(n, n)
        d[f"info{n}_x"] = xr
        tdr, ff = tfr_info.shannon_tdr_fft(xr)
        for tag, obj in (("tdr", tdr), ("fft", ff)):
            for attr in ("sig", "marginal", "info", "entropy", "isnr", "esnr"):
                d[f"info{n}_{tag}_{attr}"] = getattr(obj, attr)
            d[f"info{n}_{tag}_ref"] = np.array(obj.ref_entropy)
        d[f"info{n}_fft_angle"], d[f"info{n}_fft_freq"] = ff.angle_rads, ff.frequency
    np.savez_compressed(os.path.join(OUT, "extra.npz"), **d)
    print("extra.npz", os.path.getsize(os.path.join(OUT, "extra.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
