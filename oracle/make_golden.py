"""
make_golden.py -- generates tests/golden/*.npz by running the UNMODIFIED reference
(ISLA-UH/quantum-inferno v1.1.3, imported from /root/reference) on seeded inputs.

Run in the build container only (the reference does not travel to the GPU box):
    python oracle/make_golden.py
The fixtures pin oracle/qi_oracle.py (tests/test_oracle_golden.py) and are the golden vectors the
CUDA path is compared with on the GPU (tests/test_gpu_parity.py).  Test infrastructure.
"""
import os
import sys

import numpy as np

REF = os.environ.get("QI_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
from quantum_inferno import cwt_atoms, scales_dyadic, styx_cwt, styx_fft, styx_stx, tfr_info  # noqa: E402
from quantum_inferno.synth import benchmark_signals  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
FS = 800.0


def test_signal(n, seed=1234):
    """tone (reference synth) + linear chirp + seeded noise, n = 2^m samples at 800 Hz."""
    tone = benchmark_signals.well_tempered_tone(frequency_sample_rate_hz=FS, frequency_center_hz=60.0,
                                                time_duration_s=n / FS, time_fft_s=min(0.64, n / FS / 4))[0]
    assert len(tone) == n, (len(tone), n)
    k = np.arange(n)
    chirp = 0.5 * np.cos(2 * np.pi * (1.0 * k / FS + 0.5 * (199.0 / (n / FS)) * (k / FS) ** 2))
    noise = np.random.default_rng(seed).standard_normal(n) * 2.0 ** -4
    return tone + chirp + noise


def main():
    os.makedirs(OUT, exist_ok=True)

    # ---------------------------------------------------------------- band tables
    d = {}
    cases = []
    for fs in (100.0, 800.0, 8000.0, 48000.0, 1.0, 30.0):
        for logn in (8, 10, 13, 16, 18, 20, 22, 24, 28):
            for order in (1.0, 3.0, 6.0, 12.0, 24.0, 0.75, 0.5, 1.5):
                f = scales_dyadic.log_frequency_hz_from_fft_points(fs, 2 ** logn, scale_order=order)
                cases.append((fs, logn, order, len(f)))
                d[f"f_{len(cases) - 1}"] = f
    d["cases"] = np.array(cases)
    # the reference's own (commented-out) KAT: quantum_inferno/tests/test_scales_dyadic.py:8-21
    kat = scales_dyadic.log_frequency_hz_from_fft_points(100.0, 8192, scale_order=6.0, scale_ref_hz=1.0,
                                                         scale_base=scales_dyadic.Slice.G3)
    d["kat_fs100_n8192_o6"] = kat
    # G2 band tables used by cwt_atoms / tfr_stx_fft(is_inferno)
    bcases = []
    for (order, base, ref, lo, hi, fs) in [(3, 2.0, 1.0, 0.5, 400.0, 800.0), (6, 2.0, 1.0, 2.0, 50.0, 100.0),
                                            (12, scales_dyadic.Slice.G3, 1.0, 1.0, 24000.0, 48000.0),
                                            (1, 2.0, 1.0, 0.01, 4.0, 8.0), (3, 2.0, 1.0, 5.617, 400.0, 800.0)]:
        r = scales_dyadic.band_frequency_low_high(order, base, ref, lo, hi, fs)
        i = len(bcases)
        bcases.append((order, base, ref, lo, hi, fs))
        d[f"b{i}_band"] = r[2]
        d[f"b{i}_calg"] = r[4]
        d[f"b{i}_cgeo"] = r[5]
        d[f"b{i}_start"] = r[6]
        d[f"b{i}_end"] = r[7]
    d["bcases"] = np.array(bcases)
    np.savez_compressed(os.path.join(OUT, "scales.npz"), **d)

    # ---------------------------------------------------------------- CWT (styx_cwt)
    d = {}
    x2048 = test_signal(2048)
    x1024 = test_signal(1024, seed=99)
    d["x2048"], d["x1024"] = x2048, x1024
    for tag, x, order, dic in [("n2048_o3_norm", x2048, 3, "norm"), ("n1024_o3_spect", x1024, 3, "spect"),
                               ("n1024_o3_unit", x1024, 3, "unit"), ("n1024_o6_norm", x1024, 6, "norm"),
                               ("n1024_o12_norm", x1024, 12, "norm"), ("n1024_o1_norm", x1024, 1, "norm")]:
        f, t, c = styx_cwt.cwt_complex_any_scale_pow2(order, x, FS, cwt_type="fft", dictionary_type=dic)
        d[f"{tag}_f"], d[f"{tag}_c"] = f, c
    atoms, t_s, scale, omega, amp = styx_cwt.wavelet_centered_4cwt(3, 512, np.array([5.0, 50.0, 200.0]), FS, "norm")
    d["atoms512"], d["atoms512_t"], d["atoms512_scale"], d["atoms512_omega"], d["atoms512_amp"] = \
        atoms, t_s, scale[:, 0], omega[:, 0], amp[:, 0]
    # survey KATs on the reference tone (8192 @ 800 Hz)
    tone = benchmark_signals.well_tempered_tone()[0]
    f, t, c = styx_cwt.cwt_complex_any_scale_pow2(3, tone, FS)
    p = np.abs(c) ** 2
    g = tfr_info.shannon_stft_from_tfr_power(p)
    pf = tfr_info.ShannonStftPerFreq(p)
    pt = tfr_info.ShannonStftPerTime(p)
    bits = tfr_info.power_dynamics_scaled_bits(p)
    d["tone8192"] = tone
    d["kat8192_f"] = f
    d["kat8192_band_sum"] = p.sum(axis=1)
    d["kat8192_scalars"] = np.array([p.sum(), p.max(), g.shannon_bits.sum(), g.ref_bits, g.isnr.max(), bits[0].min()])
    d["kat8192_band_entropy"] = pf.shannon_bits.sum(axis=1)
    d["kat8192_time_entropy"] = pt.shannon_bits.sum(axis=0)
    d["kat8192_row19"] = c[19]
    d["kat8192_row0"] = c[0]
    d["kat8192_row26"] = c[26]
    np.savez_compressed(os.path.join(OUT, "cwt.npz"), **d)

    # ---------------------------------------------------------------- STX (styx_stx)
    d = {"x2048": x2048, "x1024": x1024}
    for tag, x, order in [("n2048_o3", x2048, 3), ("n1024_o6", x1024, 6), ("n1024_o12", x1024, 12)]:
        f, t, c = styx_stx.stx_complex_any_scale_pow2(order, x, FS)
        d[f"{tag}_f"], d[f"{tag}_c"] = f, c
    x256 = test_signal(256, seed=5)
    d["x256"] = x256
    for tag, kw in [("lin", dict()), ("geo", dict(is_geometric=True)), ("inf", dict(is_geometric=True, is_inferno=True)),
                    ("opt", dict(factor_q=0.5, power_p=1.0, power_r=0.5, frequency_min=10.0, frequency_max=300.0,
                                 frequency_step=5.0))]:
        r = styx_stx.tfr_stx_fft(x256, 1 / FS, scale_order_input=3.0, n_fft_in=256, **kw)
        for name, arr in zip(("tfr", "psd", "f", "ffft", "win"), r):
            d[f"gen_{tag}_{name}"] = arr
    np.savez_compressed(os.path.join(OUT, "stx.npz"), **d)

    # ---------------------------------------------------------------- STFT (styx_fft)
    d = {"tone8192": tone}
    xb = np.stack([test_signal(4096, seed=s) for s in (1, 2, 3)])
    d["xb"] = xb
    f, t, z = styx_fft.stft_complex_pow2(tone, FS, 1024, alpha=1.0)
    d["hann_f"], d["hann_t"], d["hann_z"] = f, t, z
    f, t, z = styx_fft.stft_complex_pow2(xb, FS, 256)
    d["tukey_f"], d["tukey_t"], d["tukey_z"] = f, t, z
    f, t, z = styx_fft.stft_complex_pow2(xb[0], FS, 200, overlap_points=150, nfft_points=512, alpha=0.5)
    d["odd_f"], d["odd_t"], d["odd_z"] = f, t, z
    f, t, z = styx_fft.gtx_complex_pow2(xb, FS, 512)
    d["gtx_f"], d["gtx_t"], d["gtx_z"] = f, t, z
    f, pw = styx_fft.welch_power_pow2(xb, FS, 512)
    d["welch_f"], d["welch_p"] = f, pw
    z, zb, t, f = styx_fft.stft_from_sig(tone, FS, 3)
    d["sfs_z"], d["sfs_bits"], d["sfs_t"], d["sfs_f"] = z, zb, t, f
    np.savez_compressed(os.path.join(OUT, "stft.npz"), **d)

    # ---------------------------------------------------------------- cwt_atoms
    d = {"x1024": x1024}
    for tag, kw in [("fft_norm", dict(cwt_type="fft")), ("conv_norm", dict(cwt_type="conv")),
                    ("fft_spect", dict(cwt_type="fft", dictionary_type="spect")),
                    ("fft_shift", dict(cwt_type="fft", index_shift=1.0)),
                    ("fft_o6", dict(cwt_type="fft", band_order_nth=6))]:
        c, cb, t, f = cwt_atoms.cwt_chirp_from_sig(x1024, FS, **kw)
        d[f"{tag}_c"], d[f"{tag}_bits"], d[f"{tag}_f"] = c, cb, f
    d["spec"], d["spec_f"] = cwt_atoms.chirp_spectrum(np.linspace(0.0, 400.0, 64), 0.1, 3, 50.0, FS, 1.0)
    d["spec_c"], d["spec_c_f"] = cwt_atoms.chirp_spectrum_centered(6, 80.0, FS, -1.0, scales_dyadic.Slice.G3)
    d["mqg"] = np.array([cwt_atoms.chirp_mqg_from_n(n_, s_, b_) for n_, s_, b_ in
                         [(3, 0, 2.0), (12, 1.0, 2.0), (6, -1.0, scales_dyadic.Slice.G3), (1, 0, 2.0)]])
    np.savez_compressed(os.path.join(OUT, "atoms.npz"), **d)

    # ---------------------------------------------------------------- tfr_info
    f, t, c = styx_cwt.cwt_complex_any_scale_pow2(3, x1024, FS)
    p = np.abs(c) ** 2
    d = {"power": p}
    for tag, obj in [("glob", tfr_info.shannon_stft_from_tfr_power(p)), ("ptime", tfr_info.ShannonStftPerTime(p)),
                     ("pfreq", tfr_info.ShannonStftPerFreq(p))]:
        d[f"{tag}_info"], d[f"{tag}_bits"], d[f"{tag}_isnr"], d[f"{tag}_esnr"] = obj.info, obj.shannon_bits, obj.isnr, obj.esnr
        d[f"{tag}_ref"] = np.array(obj.ref_bits)
    b0, b1, b2 = tfr_info.power_dynamics_scaled_bits(p)
    d["dyn_bits"], d["dyn_time"], d["dyn_freq"] = b0, b1, b2
    tdr, fft_ = tfr_info.shannon_tdr_fft(x1024)
    for tag, obj in [("tdr", tdr), ("fft", fft_)]:
        d[f"{tag}_marg"], d[f"{tag}_info"], d[f"{tag}_ent"], d[f"{tag}_isnr"], d[f"{tag}_esnr"] = \
            obj.marginal, obj.info, obj.entropy, obj.isnr, obj.esnr
        d[f"{tag}_ref"] = np.array(obj.ref_entropy)
    d["fft_angle"] = fft_.angle_rads
    d["fft_sig"] = fft_.sig
    d["x1024"] = x1024
    np.savez_compressed(os.path.join(OUT, "info.npz"), **d)

    # ---------------------------------------------------------------- utilities/short_time_fft
    from quantum_inferno.utilities import short_time_fft
    tone, _, fft_nd, fs_t, _, _ = benchmark_signals.well_tempered_tone()
    xs = tone + np.random.default_rng(77).standard_normal(len(tone)) / 8.0
    d = {"x": xs, "fft_nd": np.array(fft_nd)}
    cases = [(fft_nd, fft_nd // 2, "magnitude", "zeros", 0.25), (200, 150, "psd", "even", 0.5),
             (256, 192, "none", "odd", 1.0), (128, 64, "magnitude", "edge", 0.0), (100, 20, "magnitude", "zeros", 0.25)]
    d["cases"] = np.array([(m, ov, al) for m, ov, _, _, al in cases])
    d["case_scaling"] = np.array([c[2] for c in cases])
    d["case_padding"] = np.array([c[3] for c in cases])
    for i, (m, ov, scal, pad, al) in enumerate(cases):
        scal = None if scal == "none" else scal
        f, t, mag = short_time_fft.stft_tukey(xs, fs_t, al, m, ov, scal, pad)
        _, _, sp = short_time_fft.spectrogram_tukey(xs, fs_t, al, m, ov, scal, pad)
        spec = short_time_fft.get_stft_object_tukey(fs_t, al, m, ov, scal).stft(xs)
        ts, xr = short_time_fft.istft_tukey(spec, fs_t, al, m, ov, scal)
        d[f"c{i}_f"], d[f"c{i}_t"], d[f"c{i}_mag"], d[f"c{i}_sp"] = f, t, mag, sp
        d[f"c{i}_spec"], d[f"c{i}_ts"], d[f"c{i}_xr"] = spec, ts, xr
    np.savez_compressed(os.path.join(OUT, "stft_tukey.npz"), **d)

    for fn in sorted(os.listdir(OUT)):
        print(fn, os.path.getsize(os.path.join(OUT, fn)) // 1024, "KiB")


if __name__ == "__main__":
    main()
