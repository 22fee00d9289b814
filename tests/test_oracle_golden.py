"""Pins oracle/qi_oracle.py against vectors produced by the reference itself (oracle/make_golden.py)."""
import numpy as np
import pytest

from oracle import qi_oracle as orc

FS = 800.0


def rel(a, b):
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def test_band_tables_bit_exact(golden):
    g = golden("scales")
    for i, (fs, logn, order, nb) in enumerate(g["cases"]):
        f = orc.log_frequency_hz_from_fft_points(fs, 2 ** int(logn), order)
        assert len(f) == int(nb)
        assert np.array_equal(f, g[f"f_{i}"]), (fs, logn, order)


def test_reference_commented_kat(golden):
    # quantum_inferno/tests/test_scales_dyadic.py:8-21 (commented out upstream, still exact)
    f = orc.log_frequency_hz_from_fft_points(100.0, 8192, 6.0)
    assert f[0] == 0.1778279410038923 and f[-1] == 39.810717055349706 and len(f) == 48
    assert np.array_equal(f, golden("scales")["kat_fs100_n8192_o6"])


def test_g2_band_tables(golden):
    g = golden("scales")
    for i, (order, base, ref, lo, hi, fs) in enumerate(g["bcases"]):
        r = orc.band_frequency_low_high(order, base, ref, lo, hi, fs)
        assert np.array_equal(r[2], g[f"b{i}_band"])
        assert np.array_equal(r[4], g[f"b{i}_calg"])
        assert np.array_equal(r[5], g[f"b{i}_cgeo"])
        assert np.array_equal(r[6], g[f"b{i}_start"])
        assert np.array_equal(r[7], g[f"b{i}_end"])


@pytest.mark.parametrize("tag,xkey,order,dic", [
    ("n2048_o3_norm", "x2048", 3, "norm"), ("n1024_o3_spect", "x1024", 3, "spect"),
    ("n1024_o3_unit", "x1024", 3, "unit"), ("n1024_o6_norm", "x1024", 6, "norm"),
    ("n1024_o12_norm", "x1024", 12, "norm"), ("n1024_o1_norm", "x1024", 1, "norm")])
def test_cwt(golden, tag, xkey, order, dic):
    g = golden("cwt")
    f, t, c = orc.cwt_complex_any_scale_pow2(order, g[xkey], FS, dictionary_type=dic)
    assert np.array_equal(f, g[tag + "_f"])
    assert rel(c, g[tag + "_c"]) < 1e-13


def test_atoms(golden):
    g = golden("cwt")
    for b, f in enumerate([5.0, 50.0, 200.0]):
        a, scale, omega, amp = orc.gabor_atom_centered(3, 512, f, FS)
        assert rel(a, g["atoms512"][b]) < 1e-15
        assert scale == g["atoms512_scale"][b] and omega == g["atoms512_omega"][b] and amp == g["atoms512_amp"][b]


def test_cwt_survey_kats(golden):
    g = golden("cwt")
    r = orc.cwt_power_entropy(3, g["tone8192"], FS)
    assert r["power"].shape == (27, 8192)
    s = g["kat8192_scalars"]
    # values quoted in SURVEY.md section 8(c), produced by the reference
    assert abs(s[0] - 114083.52473835417) < 1e-6 and abs(s[2] - 13.925869524312288) < 1e-10
    assert abs(r["total"] - s[0]) / s[0] < 1e-13
    assert abs(r["power"].max() - s[1]) / s[1] < 1e-13
    assert abs(r["entropy_bits"] - s[2]) < 1e-11
    assert np.max(np.abs(r["band_entropy_bits"] - g["kat8192_band_entropy"])) < 1e-11
    assert np.max(np.abs(r["band_sum"] - g["kat8192_band_sum"]) / g["kat8192_band_sum"]) < 1e-12


@pytest.mark.parametrize("tag,xkey,order", [("n2048_o3", "x2048", 3), ("n1024_o6", "x1024", 6), ("n1024_o12", "x1024", 12)])
def test_stx(golden, tag, xkey, order):
    g = golden("stx")
    f, t, c = orc.stx_complex_any_scale_pow2(order, g[xkey], FS)
    assert np.array_equal(f, g[tag + "_f"])
    assert rel(c, g[tag + "_c"]) < 1e-13


@pytest.mark.parametrize("tag,kw", [
    ("lin", dict()), ("geo", dict(is_geometric=True)), ("inf", dict(is_geometric=True, is_inferno=True)),
    ("opt", dict(factor_q=0.5, power_p=1.0, power_r=0.5, frequency_min=10.0, frequency_max=300.0, frequency_step=5.0))])
def test_stx_general(golden, tag, kw):
    g = golden("stx")
    tfr, psd, f, ffft, win = orc.tfr_stx_fft(g["x256"], 1 / FS, order=3.0, n_fft_in=256, **kw)
    assert np.array_equal(f, g[f"gen_{tag}_f"]) and np.array_equal(ffft, g[f"gen_{tag}_ffft"])
    assert rel(tfr, g[f"gen_{tag}_tfr"]) < 1e-13
    assert rel(psd, g[f"gen_{tag}_psd"]) < 1e-13
    assert rel(win, g[f"gen_{tag}_win"]) < 1e-14


def test_stft(golden):
    g = golden("stft")
    f, t, z = orc.stft_complex_pow2(g["tone8192"], FS, 1024, alpha=1.0)
    assert np.array_equal(f, g["hann_f"]) and np.allclose(t, g["hann_t"], rtol=0, atol=1e-12)
    assert z.shape == g["hann_z"].shape == (513, 17) and rel(z, g["hann_z"]) < 1e-13
    f, t, z = orc.stft_complex_pow2(g["xb"], FS, 256)
    assert z.shape == g["tukey_z"].shape and rel(z, g["tukey_z"]) < 1e-13
    f, t, z = orc.stft_complex_pow2(g["xb"][0], FS, 200, overlap_points=150, nfft_points=512, alpha=0.5)
    assert z.shape == g["odd_z"].shape and rel(z, g["odd_z"]) < 1e-13
    assert np.allclose(t, g["odd_t"], rtol=0, atol=1e-12)
    f, t, z = orc.gtx_complex_pow2(g["xb"], FS, 512)
    assert z.shape == g["gtx_z"].shape and rel(z, g["gtx_z"]) < 1e-13
    f, p = orc.welch_power_pow2(g["xb"], FS, 512)
    assert np.array_equal(f, g["welch_f"]) and rel(p, g["welch_p"]) < 1e-13
    z, zb, t, f = orc.stft_from_sig(g["tone8192"], FS, 3)
    assert z.shape == (257, 33) and rel(z, g["sfs_z"]) < 1e-13
    sig_cells = g["sfs_bits"] > -40.0          # below that the reference itself is rounding noise (2^-52)
    assert np.max(np.abs(zb - g["sfs_bits"])[sig_cells]) < 1e-9 and np.all(zb[~sig_cells] < -39.0)
    with pytest.raises(ValueError):
        orc.stft_from_sig(g["tone8192"][:100], FS, 3)


@pytest.mark.parametrize("tag,kw", [
    ("fft_norm", dict(cwt_type="fft")), ("conv_norm", dict(cwt_type="conv")),
    ("fft_spect", dict(cwt_type="fft", dictionary_type="spect")), ("fft_shift", dict(cwt_type="fft", index_shift=1.0)),
    ("fft_o6", dict(cwt_type="fft", order=6))])
def test_cwt_atoms(golden, tag, kw):
    g = golden("atoms")
    c, cb, t, f = orc.cwt_chirp_from_sig(g["x1024"], FS, **kw)
    assert np.array_equal(f, g[tag + "_f"])
    assert rel(c, g[tag + "_c"]) < 1e-12
    with pytest.raises(ValueError):
        orc.cwt_chirp_from_sig(g["x1024"], FS, cwt_type="nope")


def test_tfr_info(golden):
    g = golden("info")
    p = g["power"]
    for tag, obj in [("glob", orc.shannon_stft_from_tfr_power(p)), ("ptime", orc.shannon_stft_per_time(p)),
                     ("pfreq", orc.shannon_stft_per_freq(p))]:
        assert np.array_equal(obj.info, g[tag + "_info"]) and np.array_equal(obj.shannon_bits, g[tag + "_bits"])
        assert np.array_equal(obj.isnr, g[tag + "_isnr"]) and np.array_equal(obj.esnr, g[tag + "_esnr"])
        assert obj.ref_bits == float(g[tag + "_ref"])
    b0, b1, b2 = orc.power_dynamics_scaled_bits(p)
    assert np.array_equal(b0, g["dyn_bits"]) and np.array_equal(b1, g["dyn_time"]) and np.array_equal(b2, g["dyn_freq"])
    tdr, ff = orc.shannon_tdr(g["x1024"]), orc.shannon_fft(g["x1024"])
    assert np.array_equal(tdr.info, g["tdr_info"]) and np.array_equal(tdr.entropy, g["tdr_ent"])
    assert rel(ff.sig, g["fft_sig"]) < 1e-14 and rel(ff.marginal, g["fft_marg"]) < 1e-13
    assert np.max(np.abs(ff.info - g["fft_info"])) < 1e-10
    assert np.max(np.abs(ff.angle_rads - g["fft_angle"])) < 1e-9


def test_short_time_fft_tukey(golden):
    """utilities/short_time_fft.py: restatement of scipy's ShortTimeFFT vs the reference's outputs,
    all four padding modes, the three scalings, odd hop/window combinations, and the reference's own round-trip test
    (quantum_inferno/tests/utilities/test_short_time_fft.py:47-66, atol 1e-14)."""
    g = golden("stft_tukey")
    x = g["x"]
    for i, (m, ov, alpha) in enumerate(g["cases"]):
        scal = None if str(g["case_scaling"][i]) == "none" else str(g["case_scaling"][i])
        pad = str(g["case_padding"][i])
        f, t, mag = orc.stft_tukey(x, FS, alpha, int(m), int(ov), scal, pad)
        assert np.array_equal(f, g[f"c{i}_f"]) and np.array_equal(t, g[f"c{i}_t"])
        assert mag.shape == g[f"c{i}_mag"].shape and rel(mag, g[f"c{i}_mag"]) < 1e-13
        _, _, sp = orc.spectrogram_tukey(x, FS, alpha, int(m), int(ov), scal, pad)
        assert rel(sp, g[f"c{i}_sp"]) < 1e-13
        ts, xr = orc.istft_tukey(g[f"c{i}_spec"], FS, alpha, int(m), int(ov), scal)
        assert np.array_equal(ts, g[f"c{i}_ts"]) and xr.shape == g[f"c{i}_xr"].shape
        assert np.max(np.abs(xr - g[f"c{i}_xr"])) < 1e-14
    # the reference's shape assertions (test_short_time_fft.py:35-45) and round trip on the first case
    nd = int(g["fft_nd"])
    assert g["c0_mag"].shape == (nd // 2 + 1, len(x) // (nd // 2) + 1)
    assert np.allclose(x[:len(g["c0_xr"])], g["c0_xr"], atol=1e-14)


# ----------------------------------------------------------------------------- after the path
def test_subsample_oracle(golden):
    g = golden("pick")
    for dt in ("float64", "float32"):
        plane = g["plane"].astype(dt)
        for f in g["factors"]:
            for m in ("average", "median", "max", "min", "nth"):
                got, want = orc.subsample_2d(plane, int(f), m), g[f"sub2d_{dt}_{int(f)}_{m}"]
                assert got.shape == want.shape and got.dtype == want.dtype
                assert np.array_equal(got, want, equal_nan=True), (dt, f, m)
    for f in (2, 5, 64, 200):
        for m in ("average", "median", "max", "min", "nth"):
            y, rate = orc.subsample(g["row"], 800.0, f, m)
            assert np.array_equal(y, g[f"sub1d_{f}_{m}"]) and rate == float(g[f"sub1d_{f}_{m}_rate"])


def test_picker_oracle(golden):
    g = golden("pick")
    x = g["x"]
    for et in ("sigmax", "sigmin", "sigabs", "log2", "log2max"):
        assert np.array_equal(orc.scale_signal_by_extraction_type(x, et), g[f"scaled_{et}"])
        for h in (0.7, 0.3):
            assert np.array_equal(orc.find_peaks_by_extraction_type(x, et, h), g[f"peaks_{et}_{h}"]), (et, h)
    for st in ("amplitude", "log2"):
        for tb in (1, 3):
            for dist in (0.1, 0.01, 0.5):
                assert np.array_equal(orc.find_peaks_with_bits(x, 800.0, st, tb, dist), g[f"bits_{st}_{tb}_{dist}"])
    assert np.array_equal(orc.find_peaks_with_bits(g["noise"], 800.0, "log2", 2, 0.05), g["noise_peaks_bits"])
    assert np.array_equal(orc.find_peaks_by_extraction_type(g["noise"], "sigmax", 0.5), g["noise_peaks_sigmax"])
    assert len(g["peaks_sigmax_0.3"]) > len(g["peaks_sigmax_0.7"]) > 0          # the fixture is not vacuous
    assert len(g["bits_log2_3_0.01"]) > len(g["bits_log2_3_0.5"]) > 0


# ----------------------------------------------------------------------------- before the path
def test_filtfilt_oracle(golden):
    """The float64 restatement of scipy's filtfilt / sosfiltfilt reproduces the reference's outputs (same recursion,
    same order of operations); the long-double run of it stays within the recursion's own rounding noise."""
    g = golden("iir")
    x = g["x"]
    for name in ("lowpass_100", "highpass_5", "bandpass_10_100", "bandpass_58_62", "bandpass_nyq", "lowpass_o5", "antialias"):
        alpha = float(g[f"{name}_alpha"])
        xt = x if alpha < 0 else x * orc.tukey(len(x), alpha)
        got = orc.filtfilt(g[f"{name}_b"], g[f"{name}_a"], xt)
        ref = g[f"{name}_ref"]
        noise = np.max(np.abs(ref - g[f"{name}_truth"])) / np.max(np.abs(ref))
        assert np.max(np.abs(got - ref)) / np.max(np.abs(ref)) <= max(2.0 * noise, 1e-14), name
        assert noise < 1e-7
    for name in ("pick_100_200_o7", "pick_1_10_o4", "pick_50_70_o2"):
        got, ref = orc.sosfiltfilt(g[f"{name}_sos"], x), g[f"{name}_ref"]
        noise = np.max(np.abs(ref - g[f"{name}_truth"])) / np.max(np.abs(ref))
        assert np.max(np.abs(got - ref)) / np.max(np.abs(ref)) <= max(2.0 * noise, 1e-14) and noise < 1e-11, name
    b, a = g["bandpass_10_100_b"], g["bandpass_10_100_a"]
    from scipy import signal
    assert np.allclose(orc.lfilter_zi(b, a), signal.lfilter_zi(b, a), rtol=1e-9, atol=0)
    assert np.allclose(orc.sosfilt_zi(g["pick_1_10_o4_sos"]), signal.sosfilt_zi(g["pick_1_10_o4_sos"]), rtol=1e-9)
    for m, al in ((100, 0.3), (101, 0.5), (64, 1.0), (10, 0.0), (7, 0.9)):
        assert np.allclose(orc.tukey(m, al), signal.windows.tukey(m, al), rtol=0, atol=1e-15)


# ----------------------------------------------------------------------------- synthetic inputs
def test_synth_oracle(golden):
    g = golden("synth")
    sig, t, nfft, fs, fc, df = orc.well_tempered_tone()
    assert np.array_equal(sig, g["tone0_sig"]) and np.array_equal(t, g["tone0_t"])
    assert np.array_equal([nfft, fs, fc, df], g["tone0_meta"])
    sig = orc.well_tempered_tone(800.0, 61.3, 20.48, 1.28, use_fft_frequency=False)[0]
    assert np.array_equal(sig, g["tone1_sig"])
    for i, kw in enumerate([dict(omega=2 * np.pi * 60 / 800, order=3), dict(omega=0.3, order=12, gamma=0.7),
                            dict(omega=0.9 * np.pi, order=6, gauss=False), dict(omega=0.2, order=3, oversample_scale=4)]):
        wf, support = orc.quantum_chirp(**kw)
        assert support == int(g[f"chirp{i}_support"])
        assert np.max(np.abs(wf - g[f"chirp{i}_wf"])) / np.max(np.abs(g[f"chirp{i}_wf"])) < 1e-12
    for q in (2, 4, 10):
        assert np.max(np.abs(orc.decimate(g["x"], q) - g[f"dec_{q}"])) / np.max(np.abs(g[f"dec_{q}"])) < 1e-11
