"""CPU-side check of the kernel LOGIC and of the Python drivers: the package's public API is run with the
test-only emulator runtime (tests/emul/, a g++ build of the same .cu sources) and compared with the golden
vectors the reference produced.  The GPU parity tests proper are in test_gpu_parity.py (-m gpu)."""
import numpy as np
import pytest

from quantum_inferno_b200 import _runtime, cwt_atoms, styx_cwt, styx_fft, styx_stx, tfr_info
from tests.emul.emul_runtime import EmulRuntime

FS = 800.0


@pytest.fixture(scope="module", autouse=True)
def emul():
    with _runtime.use_runtime(EmulRuntime()) as rt:
        yield rt


def rel(a, b):
    return float(np.max(np.abs(np.asarray(a) - b)) / np.max(np.abs(b)))


@pytest.mark.parametrize("tag,xkey,order,dic", [
    ("n2048_o3_norm", "x2048", 3, "norm"), ("n1024_o3_spect", "x1024", 3, "spect"),
    ("n1024_o3_unit", "x1024", 3, "unit"), ("n1024_o6_norm", "x1024", 6, "norm"),
    ("n1024_o12_norm", "x1024", 12, "norm"), ("n1024_o1_norm", "x1024", 1, "norm")])
@pytest.mark.parametrize("dtype,tol", [("float64", 1e-10), ("float32", 2e-6)])
def test_cwt(golden, tag, xkey, order, dic, dtype, tol):
    g = golden("cwt")
    f, t, c = styx_cwt.cwt_complex_any_scale_pow2(order, g[xkey], FS, dictionary_type=dic, dtype=dtype)
    assert np.array_equal(f, g[tag + "_f"])
    assert c.shape == g[tag + "_c"].shape and rel(c, g[tag + "_c"]) < tol
    assert np.array_equal(t, np.arange(len(g[xkey])) / FS)


def test_cwt_table_spectrum_and_batch(golden):
    g = golden("cwt")
    x = np.stack([g["x1024"], g["x1024"][::-1]])
    f, t, c = styx_cwt.cwt_complex_any_scale_pow2(3, x, FS, spectrum="table")
    assert c.shape == (2, 18, 1024)
    assert rel(c[0], g["n1024_o3_norm_c"] if "n1024_o3_norm_c" in g.files else c[0]) < 1e-12
    f2, _, (c2, p2) = styx_cwt.cwt_complex_any_scale_pow2(3, x, FS, outputs="both")
    assert rel(c2, c) < 1e-12 and rel(p2, np.abs(c) ** 2) < 1e-12
    with pytest.raises(NotImplementedError):
        styx_cwt.cwt_complex_any_scale_pow2(3, g["x1024"], FS, cwt_type="morlet2")


def test_atoms(golden):
    g = golden("cwt")
    atoms, t_s, scale, omega, amp = styx_cwt.wavelet_centered_4cwt(3, 512, np.array([5.0, 50.0, 200.0]), FS, "norm")
    assert rel(atoms, g["atoms512"]) < 1e-13
    assert np.array_equal(t_s, g["atoms512_t"])
    assert np.array_equal(scale[:, 0], g["atoms512_scale"]) and scale.shape == (3, 512)
    assert np.array_equal(omega[:, 0], g["atoms512_omega"]) and np.array_equal(amp[:, 0], g["atoms512_amp"])
    a1, _, s1, o1, m1 = styx_cwt.wavelet_centered_4cwt(3, 512, 50.0, FS)
    assert a1.shape == (512,) and rel(a1, g["atoms512"][1]) < 1e-13 and s1 == g["atoms512_scale"][1]


@pytest.mark.parametrize("tag,xkey,order", [("n2048_o3", "x2048", 3), ("n1024_o6", "x1024", 6), ("n1024_o12", "x1024", 12)])
@pytest.mark.parametrize("dtype,tol", [("float64", 1e-10), ("float32", 2e-6)])
def test_stx(golden, tag, xkey, order, dtype, tol):
    g = golden("stx")
    f, t, c = styx_stx.stx_complex_any_scale_pow2(order, g[xkey], FS, dtype=dtype)
    assert np.array_equal(f, g[tag + "_f"]) and rel(c, g[tag + "_c"]) < tol


def test_stft_interior_fused_path():
    """Records long enough for interior CTAs (frame means from hop-block sums, (x - mean) * window applied in the
    gather): every hop ratio of the fused path and one that falls back, against the oracle's scipy restatement."""
    from oracle import qi_oracle as orc
    k = np.arange(40000)
    x = np.random.default_rng(0).standard_normal((2, 40000)) + 3.0 + np.cos(2 * np.pi * 60 / 800 * k)
    # 512 / 1024 / 2048-point frames take the compile-time instantiations (register gather), the others the generic kernel
    for seg, ov, nfft in ((256, None, None), (256, 192, 512), (128, 96, None), (300, 100, 512), (200, 150, 256), (250, 100, 256),
                          (1024, None, None), (700, 300, 1024), (2048, None, None), (1500, 1100, 2048), (512, 128, None)):
        f0, t0, z0 = orc.stft_complex_pow2(x, FS, seg, ov, nfft, alpha=0.25)
        for dtype, tol in (("float64", 1e-12), ("float32", 3e-6)):
            f, t, z = styx_fft.stft_complex_pow2(x, FS, seg, ov, nfft, alpha=0.25, dtype=dtype)
            assert z.shape == z0.shape and rel(z, z0) < tol, (seg, ov, nfft, dtype)
    f, p = styx_fft.welch_power_pow2(x, FS, 256)
    assert rel(p, orc.welch_power_pow2(x, FS, 256)[1]) < 1e-12


@pytest.mark.parametrize("tag,kw", [
    ("lin", dict()), ("geo", dict(is_geometric=True)), ("inf", dict(is_geometric=True, is_inferno=True)),
    ("opt", dict(factor_q=0.5, power_p=1.0, power_r=0.5, frequency_min=10.0, frequency_max=300.0, frequency_step=5.0))])
def test_stx_general(golden, tag, kw):
    g = golden("stx")
    tfr, psd, f, ffft, win = styx_stx.tfr_stx_fft(g["x256"], 1 / FS, scale_order_input=3.0, n_fft_in=256, **kw)
    assert np.array_equal(f, g[f"gen_{tag}_f"]) and np.array_equal(ffft, g[f"gen_{tag}_ffft"])
    assert rel(tfr, g[f"gen_{tag}_tfr"]) < 1e-10 and rel(psd, g[f"gen_{tag}_psd"]) < 1e-10
    assert rel(win, g[f"gen_{tag}_win"]) < 1e-12


def test_stx_general_many_bands():
    """n_fft >= 512 on the default linear grid has more than 170 bands: the window workspace is sized by the library
    (round-1 regression: the driver reserved 16 bytes per band, the kernel needs 24)."""
    from oracle import qi_oracle as orc
    x = np.random.default_rng(7).standard_normal(1024) + np.cos(2 * np.pi * 60 / FS * np.arange(1024))
    tfr, psd, f, ffft, win = styx_stx.tfr_stx_fft(x, 1 / FS, scale_order_input=3.0, n_fft_in=1024)
    tfr0, psd0, f0, ffft0, win0 = orc.tfr_stx_fft(x, 1 / FS, order=3.0, n_fft_in=1024)
    assert len(f) > 170 and np.array_equal(f, f0) and np.array_equal(ffft, ffft0)
    assert rel(tfr, tfr0) < 1e-10 and rel(psd, psd0) < 1e-10 and rel(win, win0) < 1e-12


def test_stx_multirate():
    """Decimated voices + polyphase interpolation (float32 opt-in) against the oracle: north-star float32 tolerance
    1e-4 relative L2 per plane, and per band (no band may hide behind the strong ones)."""
    from oracle import qi_oracle as orc
    for n, order in ((8192, 3), (4096, 12)):
        k = np.arange(n)
        x = np.random.default_rng(n).standard_normal((2, n)) / 4 + np.cos(2 * np.pi * 60 / FS * k)
        f, t, c = styx_stx.stx_complex_any_scale_pow2(order, x, FS, dtype="float32", method="multirate")
        _, _, p = styx_stx.stx_complex_any_scale_pow2(order, x, FS, dtype="float32", method="multirate", outputs="power")
        for ch in range(2):
            f0, t0, c0 = orc.stx_complex_any_scale_pow2(order, x[ch], FS)
            assert np.array_equal(f, f0) and c.shape == (2,) + c0.shape and c.dtype == np.complex64
            assert np.linalg.norm(c[ch] - c0) / np.linalg.norm(c0) < 1e-5
            assert max(np.linalg.norm(c[ch, b] - c0[b]) / np.linalg.norm(c0[b]) for b in range(len(f))) < 2e-5
            assert np.linalg.norm(p[ch] - np.abs(c0) ** 2) / np.linalg.norm(np.abs(c0) ** 2) < 2e-5
    with pytest.raises(ValueError):
        styx_stx.stx_complex_any_scale_pow2(3, np.zeros(8192), FS, method="multirate")               # float64
    with pytest.raises(ValueError):
        styx_stx.stx_complex_any_scale_pow2(3, np.zeros(2048), FS, dtype="float32", method="multirate")


def test_stx_errors(golden):
    g = golden("stx")
    with pytest.raises(TypeError):
        styx_stx.tfr_stx_fft(g["x256"], 1 / FS)                      # n_fft_in=None, as upstream
    with pytest.raises(ValueError):
        styx_stx.tfr_stx_fft(g["x256"], 1 / FS, n_fft_in=128)
    with pytest.raises(TypeError):
        styx_stx.tfr_stx_fft(g["x256"], 1 / FS, n_fft_in=512)        # real padding path is broken upstream too
    with pytest.raises(ValueError):
        styx_stx.stx_complex_any_scale_pow2(3, g["x256"][:200], FS)


@pytest.mark.parametrize("dtype,tol", [("float64", 1e-12), ("float32", 1e-6)])
def test_stft(golden, dtype, tol):
    g = golden("stft")
    f, t, z = styx_fft.stft_complex_pow2(g["tone8192"], FS, 1024, alpha=1.0, dtype=dtype)
    assert np.array_equal(f, g["hann_f"]) and np.allclose(t, g["hann_t"], rtol=0, atol=1e-12)
    assert z.shape == (513, 17) and rel(z, g["hann_z"]) < tol
    f, t, z = styx_fft.stft_complex_pow2(g["xb"], FS, 256, dtype=dtype)
    assert z.shape == g["tukey_z"].shape and rel(z, g["tukey_z"]) < tol
    f, t, z = styx_fft.stft_complex_pow2(g["xb"][0], FS, 200, overlap_points=150, nfft_points=512, alpha=0.5, dtype=dtype)
    assert z.shape == g["odd_z"].shape and rel(z, g["odd_z"]) < tol and np.allclose(t, g["odd_t"], rtol=0, atol=1e-12)
    f, t, z = styx_fft.gtx_complex_pow2(g["xb"], FS, 512, dtype=dtype)
    assert rel(z, g["gtx_z"]) < tol
    f, p = styx_fft.welch_power_pow2(g["xb"], FS, 512, dtype=dtype)
    assert np.array_equal(f, g["welch_f"]) and rel(p, g["welch_p"]) < tol
    z, zb, t, f = styx_fft.stft_from_sig(g["tone8192"], FS, 3, dtype=dtype)
    assert z.shape == (257, 33) and rel(z, g["sfs_z"]) < tol
    if dtype == "float64":
        cells = g["sfs_bits"] > -40.0
        assert np.max(np.abs(zb - g["sfs_bits"])[cells]) < 1e-9
    with pytest.raises(ValueError):
        styx_fft.stft_from_sig(g["tone8192"][:100], FS, 3)


@pytest.mark.parametrize("tag,kw", [
    ("fft_norm", dict(cwt_type="fft")), ("conv_norm", dict(cwt_type="conv")),
    ("fft_spect", dict(cwt_type="fft", dictionary_type="spect")), ("fft_shift", dict(cwt_type="fft", index_shift=1.0)),
    ("fft_o6", dict(cwt_type="fft", band_order_nth=6))])
def test_cwt_atoms(golden, tag, kw):
    g = golden("atoms")
    c, cb, t, f = cwt_atoms.cwt_chirp_from_sig(g["x1024"], FS, **kw)
    assert np.array_equal(f, g[tag + "_f"])
    assert rel(c, g[tag + "_c"]) < 1e-10
    cells = g[tag + "_bits"] > -30.0
    assert np.max(np.abs(cb - g[tag + "_bits"])[cells]) < 1e-8
    with pytest.raises(ValueError):
        cwt_atoms.cwt_chirp_from_sig(g["x1024"], FS, cwt_type="nope")
    with pytest.raises(NotImplementedError):
        cwt_atoms.cwt_chirp_from_sig(g["x1024"], FS, cwt_type="morlet2")


def test_tfr_info(golden):
    g = golden("info")
    p = g["power"]
    for tag, obj, d in [("glob", tfr_info.shannon_stft_from_tfr_power(p), p.size),
                        ("ptime", tfr_info.ShannonStftPerTime(p), p.shape[0]),
                        ("pfreq", tfr_info.ShannonStftPerFreq(p), p.shape[1])]:
        assert np.max(np.abs(obj.info - g[tag + "_info"])) < 1e-10
        assert rel(obj.shannon_bits, g[tag + "_bits"]) < 1e-12
        assert np.max(np.abs(obj.isnr - g[tag + "_isnr"])) < 1e-10 and rel(obj.esnr, g[tag + "_esnr"]) < 1e-12
        assert obj.ref_bits == float(g[tag + "_ref"])
    b0, b1, b2 = tfr_info.power_dynamics_scaled_bits(p)
    assert np.max(np.abs(b0 - g["dyn_bits"])) < 1e-10 and np.max(np.abs(b1 - g["dyn_time"])) < 1e-10
    assert np.max(np.abs(b2 - g["dyn_freq"])) < 1e-10
    tdr, ff = tfr_info.shannon_tdr_fft(g["x1024"])
    assert np.max(np.abs(tdr.info - g["tdr_info"])) < 1e-10 and rel(tdr.entropy, g["tdr_ent"]) < 1e-12
    assert rel(tdr.marginal, g["tdr_marg"]) < 1e-13 and tdr.ref_entropy == float(g["tdr_ref"])
    assert rel(ff.sig, g["fft_sig"]) < 1e-13 and rel(ff.marginal, g["fft_marg"]) < 1e-12
    assert np.max(np.abs(ff.info - g["fft_info"])) < 1e-9 and np.max(np.abs(ff.angle_rads - g["fft_angle"])) < 1e-8
    # batch extension == reference applied per matrix
    pb = np.stack([p, 2.0 * p[::-1]])
    ob = tfr_info.ShannonStftPerFreq(pb)
    assert rel(ob.shannon_bits[0], g["pfreq_bits"]) < 1e-12 and ob.info.shape == pb.shape
    # float32 planes stay inside the north-star tolerance (1e-3 bits)
    o32 = tfr_info.shannon_stft_from_tfr_power(p, dtype="float32")
    assert abs(float(o32.shannon_bits.sum()) - float(g["glob_bits"].sum())) < 1e-3


def _synth(n, seed=0):
    k = np.arange(n)
    chirp = 0.5 * np.cos(2 * np.pi * (1.0 * k / FS + 0.5 * (199.0 / (n / FS)) * (k / FS) ** 2))
    return np.cos(2 * np.pi * 60.0 / FS * k) + chirp + np.random.default_rng(seed).standard_normal(n) / 16.0


@pytest.mark.parametrize("order,logn", [(3, 13), (3, 15), (6, 13), (12, 13), (1, 14)])
def test_multirate_fused_path(order, logn):
    """The fp32 fast path (pyramid + overlap-save + half-band interpolation) against the fp64 oracle, inside the
    north-star fp32 tolerance: power rel. L2 <= 1e-4, entropy <= 1e-3 bits."""
    from oracle import qi_oracle as orc
    from quantum_inferno_b200 import cwt_entropy
    x = np.stack([_synth(1 << logn, 1), _synth(1 << logn, 2)[::-1]])
    r = cwt_entropy.cwt_power_entropy(order, x, FS, dtype="float32", method="multirate")
    for c in range(2):
        ref = orc.cwt_power_entropy(order, x[c], FS)
        l2 = np.linalg.norm(r.power[c] - ref["power"]) / np.linalg.norm(ref["power"])
        assert l2 < 2e-5, l2
        assert abs(float(r.entropy_bits()[c]) - ref["entropy_bits"]) < 1e-4
        assert np.max(np.abs(r.band_power[c] - ref["band_sum"]) / ref["band_sum"].max()) < 1e-5
        assert abs(r.total_power[c] - ref["total"]) / ref["total"] < 1e-5
        # every band individually, the record-long (truncated) atoms of the lowest bands included
        per_band = np.linalg.norm(r.power[c] - ref["power"], axis=1) / np.linalg.norm(ref["power"], axis=1)
        assert per_band.max() < 5e-5, per_band.max()


def test_multirate_complex_output():
    """styx_cwt(dtype=float32, method='multirate'): complex TFR from the fast path."""
    from oracle import qi_oracle as orc
    x = _synth(8192, 4)
    fr, tr, cr = orc.cwt_complex_any_scale_pow2(3, x, FS)
    f, t, c = styx_cwt.cwt_complex_any_scale_pow2(3, x, FS, dtype="float32", method="multirate")
    assert c.dtype == np.complex64 and np.array_equal(f, fr)
    assert np.max(np.abs(c - cr)) / np.max(np.abs(cr)) < 5e-5
    assert np.linalg.norm(np.abs(c) ** 2 - np.abs(cr) ** 2) / np.linalg.norm(np.abs(cr) ** 2) < 2e-5
    f, t, (c2, p2) = styx_cwt.cwt_complex_any_scale_pow2(3, x, FS, dtype="float32", method="multirate", outputs="both")
    assert np.array_equal(c2, c) and np.allclose(p2, np.abs(c) ** 2, rtol=1e-5)
    with pytest.raises(ValueError):
        styx_cwt.cwt_complex_any_scale_pow2(3, x, FS, method="multirate")          # float64 default


def test_multirate_method_selection():
    from quantum_inferno_b200 import cwt_entropy
    x = _synth(8192)
    a = cwt_entropy.cwt_power_entropy(3, x, FS, dtype="float32")                  # auto -> multirate
    b = cwt_entropy.cwt_power_entropy(3, x, FS, dtype="float32", method="exact")
    assert np.linalg.norm(a.power - b.power) / np.linalg.norm(b.power) < 2e-5
    assert not np.array_equal(a.power, b.power)
    with pytest.raises(ValueError):
        cwt_entropy.cwt_power_entropy(3, x, FS, dtype="float64", method="multirate")
    with pytest.raises(ValueError):
        cwt_entropy.cwt_power_entropy(3, x[:1000], FS, dtype="float32", method="multirate")
    # non power-of-two records silently use the exact method under 'auto'
    c = cwt_entropy.cwt_power_entropy(3, x[:3000], FS, dtype="float32")
    assert c.power.shape == (1, len(c.frequency_hz), 3000)


@pytest.mark.parametrize("dtype,tol", [("float64", 1e-12), ("float32", 2e-5)])
def test_short_time_fft_tukey(golden, dtype, tol):
    """Drop-in utilities/short_time_fft: kernels (gather, detrend, window, FFT, phase rotation; inverse FFT pairs, dual
    window, overlap-add) against the reference's outputs."""
    from scipy.signal import ShortTimeFFT as ScipyShortTimeFFT
    from quantum_inferno_b200.utilities import short_time_fft as stf
    g = golden("stft_tukey")
    x = g["x"]
    for i, (m, ov, alpha) in enumerate(g["cases"]):
        scal = None if str(g["case_scaling"][i]) == "none" else str(g["case_scaling"][i])
        pad = str(g["case_padding"][i])
        f, t, mag = stf.stft_tukey(x, FS, alpha, int(m), int(ov), scal, pad, dtype=dtype)
        assert np.array_equal(f, g[f"c{i}_f"]) and np.array_equal(t, g[f"c{i}_t"])
        assert mag.shape == g[f"c{i}_mag"].shape and rel(mag, g[f"c{i}_mag"]) < tol
        _, _, sp = stf.spectrogram_tukey(x, FS, alpha, int(m), int(ov), scal, pad, dtype=dtype)
        assert rel(sp, g[f"c{i}_sp"]) < tol
        obj = stf.get_stft_object_tukey(FS, alpha, int(m), int(ov), scal, dtype=dtype)
        assert isinstance(obj, ScipyShortTimeFFT) and obj.invertible and obj.fft_mode == "onesided"
        assert rel(obj.stft(x), g[f"c{i}_spec"]) < tol
        ts, xr = stf.istft_tukey(g[f"c{i}_spec"], FS, alpha, int(m), int(ov), scal, dtype=dtype)
        assert np.array_equal(ts, g[f"c{i}_ts"]) and xr.shape == g[f"c{i}_xr"].shape
        assert np.max(np.abs(xr - g[f"c{i}_xr"])) < (1e-14 if dtype == "float64" else 2e-5)
    # the reference's own tests (quantum_inferno/tests/utilities/test_short_time_fft.py:17-66)
    nd = int(g["fft_nd"])
    obj = stf.get_stft_object_tukey(FS, 0.25, nd, nd // 2, "magnitude")
    assert (obj.scaling, obj.m_num, obj.mfft, obj.hop) == ("magnitude", nd, nd, nd // 2) and obj.delta_f == FS / nd
    ts, xr = stf.istft_tukey(obj.stft(x), FS, 0.25, nd, nd // 2, "magnitude")
    assert len(xr) == len(x) and np.allclose(x, xr, atol=1e-14) and np.allclose(ts, np.arange(len(x)) / FS, atol=1e-14)
    batch = np.stack([x, x[::-1]])
    assert rel(obj.stft(batch)[1], obj.stft(x[::-1].copy())) < 1e-15
    with pytest.raises(NotImplementedError):
        obj.stft_detrend(x, "linear")


# ----------------------------------------------------------------------------- after the path
def test_subsample(golden, capsys):
    from tests import _pick_checks as pc
    pc.check_subsample_golden(golden)
    pc.check_subsample_edges(capsys)
    pc.check_subsample_vs_oracle()
    pc.check_subsample_vector_path()


def test_picker(golden, capsys):
    from tests import _pick_checks as pc
    pc.check_picker_golden(golden)
    pc.check_picker_edges(capsys)
    pc.check_picker_vs_oracle(n=30000)


# ----------------------------------------------------------------------------- before the path
def test_butter_filters(golden, capsys):
    from tests import _iir_checks as ic
    ic.check_butter_golden(golden, capsys)
    ic.check_picker_bandpass_golden(golden)
    ic.check_filtfilt_vs_oracle(n=9000)


# ----------------------------------------------------------------------------- synthetic inputs
def test_synth_and_decimate(golden, capsys):
    from tests import _synth_checks as sc
    sc.check_synth_golden(golden, capsys)
    sc.check_decimate_golden(golden)
    sc.check_noise_generators_golden(golden)


@pytest.mark.parametrize("dtype,tol", [("float64", 1e-10), ("float32", 2e-5)])
def test_cwt_band_limited_routes(dtype, tol):
    """2^13 samples is the smallest record that takes the band-limited routes of the exact path (csrc/qi_cwt_fast.cuh:
    overlap-save blocks with the per-stage twiddle tables, decimated transforms + Kaiser interpolation) and a two-pass
    FFT with the factorised inter-pass twiddles: every band against the oracle."""
    from oracle import qi_oracle as orc
    from quantum_inferno_b200 import _driver, _plan, scales_dyadic as scales
    rt = _runtime.get_runtime()
    n = 1 << 13
    xh = np.cos(2 * np.pi * 60 / FS * np.arange(n)) + 0.3 * np.random.default_rng(3).standard_normal(n)
    freq = scales.log_frequency_hz_from_fft_points(FS, n, 3)
    bands, _, _, _ = _plan.gabor_bands(3, n, freq, FS, "norm", dtype)
    out = _driver.cwt_fft(rt.asarray(xh[None, :], dtype), bands, FS, dtype, want_complex=True, want_power=True,
                          want_band_sum=True, rt=rt)
    _, _, ref = orc.cwt_complex_any_scale_pow2(3, xh, FS)
    c = np.asarray(out["complex"][0])
    assert (np.max(np.abs(c - ref), axis=1) / np.max(np.abs(ref), axis=1)).max() < tol
    p = np.asarray(out["power"][0], dtype=np.float64)
    assert np.allclose(np.asarray(out["band_sum"][0]), p.sum(axis=1), rtol=1e-6 if dtype == "float32" else 1e-12)


def test_stft_segment_longer_than_record():
    """scipy.signal.stft (reference styx_fft.py:175) warns and shortens nperseg to the record; noverlap / nfft stay."""
    import warnings
    import scipy.signal
    x = np.random.default_rng(0).standard_normal(100)
    with warnings.catch_warnings(record=True) as wref:
        warnings.simplefilter("always")
        fr, tr, zr = scipy.signal.stft(x, FS, nperseg=128, noverlap=64, nfft=128, window=("tukey", 0.25),
                                       detrend="constant", boundary="zeros", padded=True)
    with warnings.catch_warnings(record=True) as wour:
        warnings.simplefilter("always")
        f, t, z = styx_fft.stft_complex_pow2(x, FS, 128)
    assert [str(m.message) for m in wour] == [str(m.message) for m in wref] and len(wour) == 1
    assert z.shape == zr.shape == (65, 4) and np.array_equal(f, fr) and np.allclose(t, tr)
    assert rel(z, zr) < 1e-12
    with pytest.raises(ValueError, match="noverlap must be less than nperseg"):
        styx_fft.stft_complex_pow2(x[:60], FS, 128)


# ----------------------------------------------------------------------------- thinly covered rows (a5, a9, a14, errors)
def test_atoms_all_dictionaries(golden):
    from tests import _extra_checks as ec
    ec.check_atoms_all_dictionaries(golden)


def test_stx_general_multipass(golden):
    from tests import _extra_checks as ec
    ec.check_stx_general_multipass(golden)


def test_shannon_1d_all_attributes(golden):
    from tests import _extra_checks as ec
    ec.check_shannon_1d_all_attributes(golden)


def test_reference_error_paths():
    from tests import _extra_checks as ec
    ec.check_reference_error_paths()


def test_stx_band_limited_routes():
    from tests import _extra_checks as ec
    ec.check_stx_band_limited_routes(13, channels=2)


def test_stx_power_entropy():
    from tests import _extra_checks as ec
    ec.check_stx_power_entropy(13)
