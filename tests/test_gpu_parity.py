"""GPU parity tests (run on the B200 with `-m gpu`).  Every test goes through the public Python API, i.e.
through the C ABI of libqi_b200.so; the oracle (oracle/qi_oracle.py) and the reference-generated golden
vectors (tests/golden) are the checkers.  Tolerances are the north-star's: fp64 <= 1e-10 relative,
fp32 <= 1e-4 relative L2 on power and <= 1e-3 bits on entropy, band/shift indices bit-exact."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

FS = 800.0
TOL64 = 1e-10
TOL32_L2 = 1e-4
TOL32_BITS = 1e-3


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available(), "these tests need the B200"
    from quantum_inferno_b200 import _runtime
    _runtime._runtime = None
    rt = _runtime.get_runtime()
    assert rt.name == "cuda"
    return torch


def rel(a, b):
    return float(np.max(np.abs(np.asarray(a) - b)) / np.max(np.abs(b)))


def l2(a, b):
    return float(np.linalg.norm(np.asarray(a, dtype=np.float64) - b) / np.linalg.norm(b))


def synth(n, seed=0, chan=0):
    k = np.arange(n)
    f_c = 60.0 * 2.0 ** ((chan % 12) / 12.0)
    chirp = 0.5 * np.cos(2 * np.pi * (1.0 * k / FS + 0.5 * (199.0 / (n / FS)) * (k / FS) ** 2))
    return np.cos(2 * np.pi * f_c / FS * k) + chirp + np.random.default_rng(1234 + seed + chan).standard_normal(n) / 16.0


# ----------------------------------------------------------------------------- golden vectors
@pytest.mark.parametrize("tag,xkey,order,dic", [
    ("n2048_o3_norm", "x2048", 3, "norm"), ("n1024_o3_spect", "x1024", 3, "spect"),
    ("n1024_o3_unit", "x1024", 3, "unit"), ("n1024_o6_norm", "x1024", 6, "norm"),
    ("n1024_o12_norm", "x1024", 12, "norm"), ("n1024_o1_norm", "x1024", 1, "norm")])
def test_cwt_golden(torch_cuda, golden, tag, xkey, order, dic):
    from quantum_inferno_b200 import styx_cwt
    g = golden("cwt")
    f, t, c = styx_cwt.cwt_complex_any_scale_pow2(order, g[xkey], FS, dictionary_type=dic)
    assert c.dtype == np.complex128 and np.array_equal(f, g[tag + "_f"])
    assert rel(c, g[tag + "_c"]) < TOL64
    f, t, c32 = styx_cwt.cwt_complex_any_scale_pow2(order, g[xkey], FS, dictionary_type=dic, dtype="float32")
    assert c32.dtype == np.complex64
    assert l2(np.abs(c32) ** 2, np.abs(g[tag + "_c"]) ** 2) < TOL32_L2


def test_cwt_survey_kats(torch_cuda, golden):
    """The reference-produced known answers quoted in SURVEY.md section 8(c) (8192-sample tone)."""
    from quantum_inferno_b200 import cwt_entropy, styx_cwt, tfr_info
    g = golden("cwt")
    x = g["tone8192"]
    f, t, c = styx_cwt.cwt_complex_any_scale_pow2(3, x, FS)
    assert c.shape == (27, 8192) and np.array_equal(f, g["kat8192_f"])
    for row in (0, 19, 26):
        assert rel(c[row], g[f"kat8192_row{row}"]) < TOL64
    p = np.abs(c) ** 2
    s = g["kat8192_scalars"]
    assert abs(p.sum() - s[0]) / s[0] < 1e-12 and abs(p.max() - s[1]) / s[1] < 1e-12
    assert np.unravel_index(p.argmax(), p.shape) == (19, 46)
    sh = tfr_info.shannon_stft_from_tfr_power(p)
    assert abs(sh.shannon_bits.sum() - s[2]) < 1e-10 and abs(sh.ref_bits - s[3]) < 1e-18
    assert abs(sh.isnr.max() - s[4]) < 1e-10
    assert abs(tfr_info.power_dynamics_scaled_bits(p)[0].min() - s[5]) < 1e-8
    assert np.max(np.abs(tfr_info.ShannonStftPerFreq(p).shannon_bits.sum(axis=1) - g["kat8192_band_entropy"])) < 1e-10
    assert np.max(np.abs(tfr_info.ShannonStftPerTime(p).shannon_bits.sum(axis=0) - g["kat8192_time_entropy"])) < 1e-10
    # fused fp32 path
    r = cwt_entropy.cwt_power_entropy(3, x, FS, dtype="float32")
    assert l2(r.power[0].double().cpu().numpy(), p) < TOL32_L2
    assert abs(float(r.entropy_bits()[0]) - s[2]) < TOL32_BITS
    assert np.max(np.abs(r.band_power[0].cpu().numpy() - g["kat8192_band_sum"]) / g["kat8192_band_sum"].max()) < 1e-5


def test_atoms_golden(torch_cuda, golden):
    from quantum_inferno_b200 import styx_cwt
    g = golden("cwt")
    atoms, t_s, scale, omega, amp = styx_cwt.wavelet_centered_4cwt(3, 512, np.array([5.0, 50.0, 200.0]), FS, "norm")
    assert rel(atoms, g["atoms512"]) < 1e-13 and np.array_equal(scale[:, 0], g["atoms512_scale"])


@pytest.mark.parametrize("tag,xkey,order", [("n2048_o3", "x2048", 3), ("n1024_o6", "x1024", 6), ("n1024_o12", "x1024", 12)])
def test_stx_golden(torch_cuda, golden, tag, xkey, order):
    from quantum_inferno_b200 import styx_stx
    g = golden("stx")
    f, t, c = styx_stx.stx_complex_any_scale_pow2(order, g[xkey], FS)
    assert np.array_equal(f, g[tag + "_f"]) and rel(c, g[tag + "_c"]) < TOL64
    f, t, c32 = styx_stx.stx_complex_any_scale_pow2(order, g[xkey], FS, dtype="float32")
    assert l2(np.abs(c32) ** 2, np.abs(g[tag + "_c"]) ** 2) < TOL32_L2


@pytest.mark.parametrize("tag,kw", [
    ("lin", dict()), ("geo", dict(is_geometric=True)), ("inf", dict(is_geometric=True, is_inferno=True)),
    ("opt", dict(factor_q=0.5, power_p=1.0, power_r=0.5, frequency_min=10.0, frequency_max=300.0, frequency_step=5.0))])
def test_stx_general_golden(torch_cuda, golden, tag, kw):
    from quantum_inferno_b200 import styx_stx
    g = golden("stx")
    tfr, psd, f, ffft, win = styx_stx.tfr_stx_fft(g["x256"], 1 / FS, scale_order_input=3.0, n_fft_in=256, **kw)
    assert np.array_equal(f, g[f"gen_{tag}_f"]) and np.array_equal(ffft, g[f"gen_{tag}_ffft"])
    assert rel(tfr, g[f"gen_{tag}_tfr"]) < TOL64 and rel(psd, g[f"gen_{tag}_psd"]) < TOL64
    assert rel(win, g[f"gen_{tag}_win"]) < 1e-12


def test_stx_general_many_bands(torch_cuda):
    """a9 at a multi-pass size: 4096-point record, default linear grid (~2000 bands), geometric and inferno grids."""
    from oracle import qi_oracle as orc
    from quantum_inferno_b200 import styx_stx
    n = 4096
    x = synth(n, chan=2)
    for kw in (dict(), dict(is_geometric=True), dict(is_geometric=True, is_inferno=True)):
        tfr, psd, f, ffft, win = styx_stx.tfr_stx_fft(x, 1 / FS, scale_order_input=3.0, n_fft_in=n, **kw)
        tfr0, psd0, f0, ffft0, win0 = orc.tfr_stx_fft(x, 1 / FS, order=3.0, n_fft_in=n, **kw)
        assert np.array_equal(f, f0) and np.array_equal(ffft, ffft0)
        assert rel(tfr, tfr0) < TOL64 and rel(psd, psd0) < TOL64 and rel(win, win0) < 1e-12
    assert len(styx_stx.tfr_stx_fft(x, 1 / FS, scale_order_input=3.0, n_fft_in=n)[2]) > 170


def test_stft_golden(torch_cuda, golden):
    from quantum_inferno_b200 import styx_fft
    g = golden("stft")
    for dtype, tol in (("float64", 1e-12), ("float32", 1e-6)):
        f, t, z = styx_fft.stft_complex_pow2(g["tone8192"], FS, 1024, alpha=1.0, dtype=dtype)
        assert z.shape == (513, 17) and rel(z, g["hann_z"]) < tol and np.array_equal(f, g["hann_f"])
        f, t, z = styx_fft.stft_complex_pow2(g["xb"], FS, 256, dtype=dtype)
        assert z.shape == g["tukey_z"].shape and rel(z, g["tukey_z"]) < tol
        f, t, z = styx_fft.stft_complex_pow2(g["xb"][0], FS, 200, overlap_points=150, nfft_points=512, alpha=0.5, dtype=dtype)
        assert rel(z, g["odd_z"]) < tol
        f, t, z = styx_fft.gtx_complex_pow2(g["xb"], FS, 512, dtype=dtype)
        assert rel(z, g["gtx_z"]) < tol
        f, p = styx_fft.welch_power_pow2(g["xb"], FS, 512, dtype=dtype)
        assert rel(p, g["welch_p"]) < tol
        z, zb, t, f = styx_fft.stft_from_sig(g["tone8192"], FS, 3, dtype=dtype)
        assert z.shape == (257, 33) and rel(z, g["sfs_z"]) < tol
    # survey KAT: sum of power 6.000001907348631, max 0.25 in bin 76 (the stationary tone ties across interior frames)
    f, t, z = styx_fft.stft_complex_pow2(g["tone8192"], FS, 1024, alpha=1.0)
    p = np.abs(z) ** 2
    assert abs(p.sum() - 6.000001907348631) < 1e-9 and abs(p.max() - 0.25) < 1e-12
    assert np.unravel_index(p.argmax(), p.shape)[0] == 76 and abs(p[76, 1] - 0.25) < 1e-12


@pytest.mark.parametrize("tag,kw", [
    ("fft_norm", dict(cwt_type="fft")), ("conv_norm", dict(cwt_type="conv")),
    ("fft_spect", dict(cwt_type="fft", dictionary_type="spect")), ("fft_shift", dict(cwt_type="fft", index_shift=1.0)),
    ("fft_o6", dict(cwt_type="fft", band_order_nth=6))])
def test_cwt_atoms_golden(torch_cuda, golden, tag, kw):
    from quantum_inferno_b200 import cwt_atoms
    g = golden("atoms")
    c, cb, t, f = cwt_atoms.cwt_chirp_from_sig(g["x1024"], FS, **kw)
    assert np.array_equal(f, g[tag + "_f"]) and rel(c, g[tag + "_c"]) < TOL64
    cells = g[tag + "_bits"] > -30.0
    assert np.max(np.abs(cb - g[tag + "_bits"])[cells]) < 1e-8


def test_tfr_info_golden(torch_cuda, golden):
    from quantum_inferno_b200 import tfr_info
    g = golden("info")
    p = g["power"]
    for tag, obj in [("glob", tfr_info.shannon_stft_from_tfr_power(p)), ("ptime", tfr_info.ShannonStftPerTime(p)),
                     ("pfreq", tfr_info.ShannonStftPerFreq(p))]:
        assert np.max(np.abs(obj.info - g[tag + "_info"])) < TOL64 and rel(obj.shannon_bits, g[tag + "_bits"]) < TOL64
        assert np.max(np.abs(obj.isnr - g[tag + "_isnr"])) < TOL64 and rel(obj.esnr, g[tag + "_esnr"]) < TOL64
    b0, b1, b2 = tfr_info.power_dynamics_scaled_bits(p)
    assert np.max(np.abs(b0 - g["dyn_bits"])) < TOL64 and np.max(np.abs(b1 - g["dyn_time"])) < TOL64
    assert np.max(np.abs(b2 - g["dyn_freq"])) < TOL64
    tdr, ff = tfr_info.shannon_tdr_fft(g["x1024"])
    assert np.max(np.abs(tdr.info - g["tdr_info"])) < TOL64 and rel(ff.marginal, g["fft_marg"]) < TOL64
    # torch in -> torch out, on the device
    pt = torch_cuda.from_numpy(p).cuda()
    o = tfr_info.ShannonStftPerFreq(pt)
    assert o.info.is_cuda and np.max(np.abs(o.info.cpu().numpy() - g["pfreq_info"])) < TOL64


# ----------------------------------------------------------------------------- live comparison with the oracle
@pytest.mark.parametrize("order,logn", [(3, 14), (3, 16), (6, 13), (12, 12)])
def test_cwt_vs_oracle(torch_cuda, order, logn):
    """config[0]-shaped case (order 3, 2^16 @ 800 Hz) and smaller ones: multi-pass FFT sizes, all bands."""
    from oracle import qi_oracle as orc
    from quantum_inferno_b200 import cwt_entropy, styx_cwt
    x = synth(1 << logn)
    fr, tr, cr = orc.cwt_complex_any_scale_pow2(order, x, FS)
    f, t, c = styx_cwt.cwt_complex_any_scale_pow2(order, x, FS)
    assert np.array_equal(f, fr) and rel(c, cr) < TOL64
    ref = orc.cwt_power_entropy(order, x, FS)
    r = cwt_entropy.cwt_power_entropy(order, x, FS, dtype="float32")
    assert l2(r.power[0].double().cpu().numpy(), ref["power"]) < TOL32_L2
    assert abs(float(r.entropy_bits()[0]) - ref["entropy_bits"]) < TOL32_BITS
    r64 = cwt_entropy.cwt_power_entropy(order, x, FS, dtype="float64")
    assert rel(r64.power[0].cpu().numpy(), ref["power"]) < TOL64
    # information of a weak cell amplifies the transform's 1e-13 absolute error by 1/P: compare where P is not tiny
    d_info = np.abs(r64.info[0].cpu().numpy() - ref["info"])
    strong = ref["power"] > 1e-4 * ref["power"].max()
    assert d_info[strong].max() < 1e-9 and d_info.max() < 1e-4
    assert abs(float(r64.entropy_bits()[0]) - ref["entropy_bits"]) < 1e-10


@pytest.mark.parametrize("order,logn", [(3, 13), (3, 16), (6, 14), (12, 13), (1, 13)])
def test_multirate_vs_oracle(torch_cuda, order, logn):
    """The fp32 fast path (method='multirate', what 'auto' picks for float32 2^m records) against the fp64 oracle."""
    from oracle import qi_oracle as orc
    from quantum_inferno_b200 import cwt_entropy
    n = 1 << logn
    x = np.stack([synth(n, chan=0), synth(n, chan=5)[::-1].copy()])
    r = cwt_entropy.cwt_power_entropy(order, x, FS, dtype="float32", method="multirate")
    for c in range(2):
        ref = orc.cwt_power_entropy(order, x[c], FS)
        p = r.power[c].double().cpu().numpy()
        assert l2(p, ref["power"]) < 2e-5                                         # tolerance is 1e-4
        assert abs(float(r.entropy_bits()[c]) - ref["entropy_bits"]) < 1e-4      # tolerance is 1e-3 bits
        assert abs(float(r.total_power[c]) - ref["total"]) / ref["total"] < 1e-5
        # every band individually (tolerance 1e-4), the record-long truncated atoms of the lowest bands included
        per_band = np.linalg.norm(p - ref["power"], axis=1) / np.linalg.norm(ref["power"], axis=1)
        assert per_band.max() < 5e-5, per_band
        # per-cell information: the fp32 amplitude error (~3e-6 of the plane maximum) is amplified by 1/|cwt|
        strong = ref["power"] > 1e-2 * ref["power"].max()
        assert np.abs(r.info[c].double().cpu().numpy() - ref["info"])[strong].max() < 1e-3


def test_multirate_properties_north_star_size(torch_cuda):
    """Size-independent checks at the bench size (2^24 samples, 60 bands): agreement of the two independent CUDA
    algorithms, sum rules, and the oracle on single bands."""
    torch = torch_cuda
    from oracle import qi_oracle as orc
    from quantum_inferno_b200 import cwt_entropy
    n = 1 << 24
    x = torch.from_numpy(synth(n, chan=3)).cuda()
    a = cwt_entropy.cwt_power_entropy(3, x, FS, dtype="float32", method="multirate")
    assert tuple(a.power.shape) == (1, 60, n)
    assert torch.allclose(a.power.double().sum(-1), a.band_power, rtol=1e-5)
    pdf_sum = (a.power.double() / a.total_power[:, None, None]).sum()
    assert abs(float(pdf_sum) - 1.0) < 2e-6
    pa = a.power[0].clone()
    ent_a = float(a.entropy_bits()[0])
    del a
    b = cwt_entropy.cwt_power_entropy(3, x, FS, dtype="float32", method="exact", want_info=True)
    assert float((pa.double() - b.power[0].double()).norm() / b.power[0].double().norm()) < 2e-5
    assert abs(ent_a - float(b.entropy_bits()[0])) < 1e-4
    xf = np.fft.fft(x.cpu().numpy().astype(np.float64), 2 * n)
    for band in (0, 1, 10, 35, 59):                      # 0 and 1: record-long truncated atoms
        row = np.abs(orc.cwt_band(xf, 3, n, b.frequency_hz[band], FS)) ** 2
        assert l2(pa[band].double().cpu().numpy(), row) < TOL32_L2


@pytest.mark.parametrize("order,logn,bands", [(6, 22, (0, 1, 2, 50, 101)), (12, 24, (0, 1, 4, 100, 214))])
def test_multirate_config_sizes_vs_oracle(torch_cuda, order, logn, bands):
    """BASELINE configs[3] (order 6, 2^22 samples, 102 bands) and the order-12 table at 2^24 samples: single bands of the
    fused fp32 path against the fp64 oracle at full size, the record-long atoms of the lowest bands included."""
    torch = torch_cuda
    from oracle import qi_oracle as orc
    from quantum_inferno_b200 import cwt_entropy
    n = 1 << logn
    xh = synth(n, chan=4)
    r = cwt_entropy.cwt_power_entropy(order, torch.from_numpy(xh).cuda(), FS, dtype="float32")
    assert r.power.shape[1] == len(orc.log_frequency_hz_from_fft_points(FS, n, order))
    xf = np.fft.fft(xh, 2 * n)
    for band in bands:
        row = np.abs(orc.cwt_band(xf, order, n, r.frequency_hz[band], FS)) ** 2
        assert l2(r.power[0, band].double().cpu().numpy(), row) < TOL32_L2, band
        assert abs(float(r.band_power[0, band]) - row.sum()) / row.sum() < TOL32_L2


@pytest.mark.parametrize("dtype,tol", [("float64", TOL64), ("float32", 2e-5)])
def test_cwt_band_limited_routes_large(torch_cuda, dtype, tol):
    """The band-limited routes of the exact path (csrc/qi_cwt_fast.cuh: overlap-save for short atoms, decimated
    transform + Kaiser interpolation for long ones) at 2^20 samples: against the plain three-pass route on every band
    (complex TFR, power, band sums) and against the oracle on single bands of each route."""
    torch = torch_cuda
    from oracle import qi_oracle as orc
    from quantum_inferno_b200 import _driver, _plan, _runtime, scales_dyadic as scales
    rt = _runtime.get_runtime()
    n = 1 << 20
    xh = np.stack([synth(n, chan=2), synth(n, chan=9)])
    x = rt.asarray(xh, dtype)
    freq = scales.log_frequency_hz_from_fft_points(FS, n, 3)
    bands, scale, _, _ = _plan.gabor_bands(3, n, freq, FS, "norm", dtype)
    fast = _driver.cwt_fft(x, bands, FS, dtype, want_complex=True, want_power=True, want_band_sum=True, rt=rt)
    plain = _driver.cwt_fft(x, bands, FS, dtype, want_complex=True, want_power=True, want_band_sum=True, rt=rt, plain_only=True)
    cf, cp = torch.view_as_real(fast["complex"]).double(), torch.view_as_real(plain["complex"]).double()
    per_band = ((cf - cp).abs().amax(dim=(2, 3)) / cp.abs().amax(dim=(2, 3))).cpu().numpy()
    assert per_band.max() < tol, per_band
    pf, pp = fast["power"].double(), plain["power"].double()
    assert float(((pf - pp).abs().amax(dim=2) / pp.amax(dim=2)).max()) < 2 * tol
    assert torch.allclose(fast["band_sum"], plain["band_sum"], rtol=2 * tol)
    xf = np.fft.fft(xh[1], 2 * n)
    for b in (0, 6, 20, 33, 40, 47):                       # table, decimated (x3), overlap-save (x2)
        row = orc.cwt_band(xf, 3, n, freq[b], FS)
        got = fast["complex"][1, b].cpu().numpy()
        # the reference's float64 time axis carries eps64 * omega * N / 2 of carrier-phase noise (4.5e-10 on the top band
        # here): the distance to the restatement is bounded by tolerance + that, the distance to the long-double-axis
        # arbiter of the same atom by the tolerance itself
        noise = orc.reference_axis_phase_noise(3, n, freq[b], FS)
        assert np.max(np.abs(got - row)) / np.max(np.abs(row)) < tol + noise, b
        arb = orc.cwt_band(xf, 3, n, freq[b], FS, arbiter=True)
        assert np.max(np.abs(got - arb)) / np.max(np.abs(arb)) < tol, b


def test_cwt_edge_cases(torch_cuda):
    from oracle import qi_oracle as orc
    from quantum_inferno_b200 import styx_cwt
    # odd / non-power-of-two record, zero record, 2-D batch, order below the 0.75 floor
    for n in (1001, 300, 257):
        x = synth(n)
        _, _, cr = orc.cwt_complex_any_scale_pow2(3, x, FS)
        _, _, c = styx_cwt.cwt_complex_any_scale_pow2(3, x, FS)
        assert c.shape == cr.shape and rel(c, cr) < TOL64
    _, _, c = styx_cwt.cwt_complex_any_scale_pow2(3, np.zeros(512), FS)
    assert not np.any(c)
    xb = np.stack([synth(2048, chan=i) for i in range(5)])
    _, _, cb = styx_cwt.cwt_complex_any_scale_pow2(3, xb, FS)
    for i in (0, 4):
        assert rel(cb[i], orc.cwt_complex_any_scale_pow2(3, xb[i], FS)[2]) < TOL64
    fr, _, cr = orc.cwt_complex_any_scale_pow2(0.5, xb[0], FS)
    f, _, c = styx_cwt.cwt_complex_any_scale_pow2(0.5, xb[0], FS)
    assert np.array_equal(f, fr) and rel(c, cr) < TOL64


def test_cwt_properties_large(torch_cuda):
    """Size-independent properties at a size the verbatim reference cannot hold (4 x 2^20, 48 bands)."""
    torch = torch_cuda
    from quantum_inferno_b200 import cwt_entropy
    n, C = 1 << 20, 4
    x = torch.from_numpy(np.stack([synth(n, chan=c) for c in range(C)])).cuda()
    r = cwt_entropy.cwt_power_entropy(3, x, FS, dtype="float32")
    assert tuple(r.power.shape) == (C, 48, n)
    # band sums == sums of the plane; pdf sums to one; entropy bounded by log2(D)
    assert torch.allclose(r.power.double().sum(-1), r.band_power, rtol=1e-6)
    pdf_sum = (r.power.double() / r.total_power[:, None, None]).sum((1, 2))
    assert torch.allclose(pdf_sum, torch.ones_like(pdf_sum), atol=2e-6)      # S is the pre-pass estimate (~1e-6)
    ent = r.entropy_bits()
    assert bool(((ent > 0) & (ent < np.log2(48 * n))).all())
    # info plane == -log2(P/S + eps) recomputed with torch in fp64
    ref_info = -torch.log2(r.power[1].double() / r.total_power[1] + np.finfo(np.float64).eps)
    assert float((r.info[1].double() - ref_info).abs().max()) < 1e-4
    # linearity: cwt(a*x0 + b*x1) power equals |a*c0 + b*c1|^2 -> check through two independent spectrum paths
    r_tab = cwt_entropy.cwt_power_entropy(3, x[:1], FS, dtype="float32", spectrum="table", band_slice=(40, 48))
    assert l2(r_tab.power[0].double().cpu().numpy(), r.power[0, 40:48].double().cpu().numpy()) < TOL32_L2
    # band sharding: two half tables + summed totals reproduce the unsharded normalisation
    tot = torch.zeros(1, dtype=torch.float64, device="cuda")
    parts = []
    for sl in ((0, 24), (24, 48)):
        part = cwt_entropy.cwt_power_entropy(3, x[:1], FS, dtype="float32", band_slice=sl, want_info=False)
        tot += part.total_power
        parts.append(part)
    assert abs(float(tot[0]) - float(r.total_power[0])) / float(r.total_power[0]) < 2e-6
    # oracle on a slab: first band rows against the CPU restatement of one band
    from oracle import qi_oracle as orc
    xf = np.fft.fft(x[0].cpu().numpy(), 2 * n)
    for b in (0, 1, 5, 20, 47):
        row = orc.cwt_band(xf, 3, n, r.frequency_hz[b], FS)
        assert l2(r.power[0, b].double().cpu().numpy(), np.abs(row) ** 2) < TOL32_L2   # b = 0, 1: record-long atoms


def test_stx_config3_slab(torch_cuda):
    """config[2] shape (2^18 samples, fp64 and fp32) on 2 channels, checked band-by-band with numpy on the host."""
    from quantum_inferno_b200 import _plan, styx_stx
    n = 1 << 18
    x = np.stack([synth(n, chan=c) for c in range(2)])
    f, t, c = styx_stx.stx_complex_any_scale_pow2(3, x, FS)
    assert c.shape == (2, 42, n)
    _, bands = _plan.stx_bands(3, n, FS)
    xf = np.fft.fft(x[1])
    w = 2 * np.pi * np.fft.fftfreq(n)
    for b in (0, 17, 41):
        ref = np.fft.ifft(np.roll(xf, -int(bands["shift"][b])) * np.exp(-0.5 * bands["sigma"][b] ** 2 * w ** 2))
        assert rel(c[1, b], ref) < TOL64
    f, t, p32 = styx_stx.stx_complex_any_scale_pow2(3, x, FS, dtype="float32", outputs="power")
    assert l2(p32[1, 17], np.abs(c[1, 17]) ** 2) < TOL32_L2


def test_stft_config2_slab(torch_cuda):
    """config[1] shape: Hann 1024 / 50 % on 2^20-sample records (4 channels), frames checked against numpy."""
    from quantum_inferno_b200 import _plan, styx_fft
    n = 1 << 20
    x = np.stack([synth(n, chan=c) for c in range(4)])
    f, t, z = styx_fft.stft_complex_pow2(x, FS, 1024, alpha=1.0)
    assert z.shape == (4, 513, 2049)
    win = _plan.periodic_window("tukey", 1.0, 1024)
    xe = np.pad(x[3], (512, 512))
    for fr in (0, 1, 1000, 2048):
        seg = xe[fr * 512: fr * 512 + 1024]
        ref = np.fft.rfft((seg - seg.mean()) * win) / win.sum()
        assert rel(z[3, :, fr], ref) < 1e-12
    f, t, z32 = styx_fft.stft_complex_pow2(x, FS, 1024, alpha=1.0, dtype="float32")
    assert l2(np.abs(z32[3]) ** 2, np.abs(z[3]) ** 2) < TOL32_L2


def test_fft_roundtrip_large(torch_cuda):
    """Three-pass FFT sizes: forward+inverse identity and Parseval at 2^22 (fp32) / 2^21 (fp64)."""
    torch = torch_cuda
    from quantum_inferno_b200 import _lib, _runtime
    rt = _runtime.get_runtime()
    for dt, cdt, code, log2n, tol in ((torch.float32, torch.complex64, 0, 22, 2e-6), (torch.float64, torch.complex128, 1, 21, 1e-13)):
        g = torch.Generator(device="cuda").manual_seed(5)
        x = torch.randn(2, 1 << log2n, 2, device="cuda", dtype=dt, generator=g)
        xc = torch.view_as_complex(x).contiguous()
        spec = torch.empty_like(xc)
        _lib.check(rt.lib, rt.lib.qi_fft_c2c(xc.data_ptr(), spec.data_ptr(), 2, log2n, 0, code, rt.stream()), "fft")
        ref = torch.fft.fft(xc[0])
        idx = torch.arange(1 << log2n, device="cuda")
        rev = torch.zeros_like(idx)
        for i in range(log2n):
            rev |= ((idx >> i) & 1) << (log2n - 1 - i)
        assert float((spec[0] - ref[rev]).abs().max() / ref.abs().max()) < tol * 10
        back = torch.empty_like(xc)
        _lib.check(rt.lib, rt.lib.qi_fft_c2c(spec.data_ptr(), back.data_ptr(), 2, log2n, 1, code, rt.stream()), "ifft")
        assert float((back - xc).abs().max() / xc.abs().max()) < tol * 10


@pytest.mark.parametrize("dtype,tol", [("float64", 1e-12), ("float32", 2e-5)])
def test_short_time_fft_tukey(torch_cuda, golden, dtype, tol):
    """utilities/short_time_fft drop-in against the reference's outputs + the reference's own
    round-trip test (quantum_inferno/tests/utilities/test_short_time_fft.py:47-66)."""
    from oracle import qi_oracle as orc
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
        assert rel(obj.stft(x), g[f"c{i}_spec"]) < tol
        ts, xr = stf.istft_tukey(g[f"c{i}_spec"], FS, alpha, int(m), int(ov), scal, dtype=dtype)
        assert np.array_equal(ts, g[f"c{i}_ts"]) and xr.shape == g[f"c{i}_xr"].shape
        assert np.max(np.abs(xr - g[f"c{i}_xr"])) < (1e-14 if dtype == "float64" else 2e-5)
    nd = int(g["fft_nd"])
    obj = stf.get_stft_object_tukey(FS, 0.25, nd, nd // 2, "magnitude")
    ts, xr = stf.istft_tukey(obj.stft(x), FS, 0.25, nd, nd // 2, "magnitude")
    assert len(xr) == len(x) and np.allclose(x, xr, atol=1e-14)
    # a larger batch against the oracle: 4 channels x 2^18 samples, CUDA tensors in -> CUDA tensors out
    torch = torch_cuda
    xb = np.stack([synth(1 << 18, chan=c) for c in range(4)])
    f, t, mag = stf.stft_tukey(torch.from_numpy(xb).cuda(), FS, 0.25, 1024, 512, dtype=dtype)
    assert mag.is_cuda and tuple(mag.shape) == (4, 513, (1 << 18) // 512 + 1)
    for c in (0, 3):
        assert rel(mag[c].double().cpu().numpy(), orc.stft_tukey(xb[c], FS, 0.25, 1024, 512)[2]) < tol
    spec = stf.get_stft_object_tukey(FS, 0.25, 1024, 512, dtype=dtype).stft(torch.from_numpy(xb).cuda())
    ts, xr = stf.istft_tukey(spec, FS, 0.25, 1024, 512, dtype=dtype)
    assert np.max(np.abs(xr.double().cpu().numpy() - xb[:, :xr.shape[-1]])) < (1e-13 if dtype == "float64" else 2e-5)


@pytest.mark.parametrize("order,logn", [(6, 20), (12, 18), (1.5, 19)])
def test_multirate_vs_exact_other_orders(torch_cuda, order, logn):
    """The envelope-decimated multirate path against the independent exact CUDA path (and the oracle on single bands)
    at sizes the oracle cannot do whole: other orders have other levels, kernel supports and band counts per level."""
    torch = torch_cuda
    from oracle import qi_oracle as orc
    from quantum_inferno_b200 import cwt_entropy
    n = 1 << logn
    x = torch.from_numpy(np.stack([synth(n, chan=1), synth(n, chan=7)])).cuda()
    a = cwt_entropy.cwt_power_entropy(order, x, FS, dtype="float32", method="multirate")
    b = cwt_entropy.cwt_power_entropy(order, x, FS, dtype="float32", method="exact", want_info=True)
    pa, pb = a.power.double(), b.power.double()
    per_band = ((pa - pb).norm(dim=-1) / pb.norm(dim=-1)).cpu().numpy()
    assert per_band.max() < TOL32_L2, per_band.max()       # every band, the record-long atoms included
    assert float((pa - pb).norm() / pb.norm()) < 2e-5
    assert torch.allclose(a.entropy_bits(), b.entropy_bits(), atol=1e-4, rtol=0)
    assert torch.allclose(a.total_power, b.total_power, rtol=1e-5)
    xf = np.fft.fft(x[0].cpu().numpy().astype(np.float64), 2 * n)
    for band in (0, 1, 3, len(a.frequency_hz) // 2, len(a.frequency_hz) - 2):
        row = np.abs(orc.cwt_band(xf, order, n, a.frequency_hz[band], FS)) ** 2
        assert l2(pa[0, band].cpu().numpy(), row) < TOL32_L2


@pytest.mark.parametrize("n_chan,chunks", [(8, 4), (5, 3), (2, 2), (3, 8)])
def test_host_pipelined_equals_resident(torch_cuda, n_chan, chunks):
    """cwt_power_entropy(host_chunks=k) on host-resident records (pinned tensor or numpy): channel groups are uploaded
    under the kernels of the previous group; channels are independent, so every output equals the resident call's."""
    torch = torch_cuda
    from quantum_inferno_b200 import cwt_entropy
    n = 1 << 15
    x = np.stack([synth(n, chan=c) for c in range(n_chan)]).astype(np.float32)
    ref = cwt_entropy.cwt_power_entropy(3, torch.from_numpy(x).cuda(), FS, dtype="float32")
    for src in (torch.from_numpy(x).pin_memory(), x):
        r = cwt_entropy.cwt_power_entropy(3, src, FS, dtype="float32", host_chunks=chunks)
        assert torch.equal(r.power, ref.power) and torch.allclose(r.info, ref.info, rtol=0, atol=1e-5)
        # the fp64 sums are accumulated with atomics: equal up to the order of the additions
        assert torch.allclose(r.band_power, ref.band_power, rtol=1e-12) and torch.allclose(r.total_power, ref.total_power, rtol=1e-12)
        assert torch.allclose(r.band_entropy_bits, ref.band_entropy_bits, rtol=0, atol=1e-9)


# ----------------------------------------------------------------------------- after the path
def test_subsample_gpu(torch_cuda, golden, capsys):
    from tests import _pick_checks as pc
    pc.check_subsample_golden(golden)
    pc.check_subsample_edges(capsys)
    pc.check_subsample_vs_oracle()
    pc.check_subsample_vector_path()


def test_picker_gpu(torch_cuda, golden, capsys):
    from tests import _pick_checks as pc
    pc.check_picker_golden(golden)
    pc.check_picker_edges(capsys)
    pc.check_picker_vs_oracle()


def test_subsample_plane_properties(torch_cuda):
    """Display reduction of a [bands, 2^22] fp32 plane kept on the device: size-independent properties
    (min <= median <= max, mean inside them, factor f then g == factor f*g for max / min / nth, tensor in -> tensor out)."""
    from quantum_inferno_b200.utilities import sampling
    torch = torch_cuda
    gen = torch.Generator(device="cuda").manual_seed(7)
    p = torch.rand((12, 1 << 22), generator=gen, device="cuda", dtype=torch.float32) ** 4
    red = {m: sampling.subsample_2d(p, 32, m) for m in ("average", "median", "max", "min", "nth")}
    assert all(isinstance(v, torch.Tensor) and v.shape == (12, 1 << 17) for v in red.values())
    assert bool((red["min"] <= red["median"]).all()) and bool((red["median"] <= red["max"]).all())
    assert bool((red["min"] <= red["average"]).all()) and bool((red["average"] <= red["max"]).all())
    assert torch.equal(red["max"], p.reshape(12, -1, 32).amax(dim=2)) and torch.equal(red["nth"], p[:, ::32])
    for m in ("max", "min", "nth"):
        assert torch.equal(sampling.subsample_2d(sampling.subsample_2d(p, 32, m), 64, m), sampling.subsample_2d(p, 2048, m))
    assert torch.equal(sampling.subsample_2d(p, 2048, "median"), p.reshape(12, -1, 2048).sort(dim=2).values[:, :, 1023:1025].sum(dim=2) / 2)


# ----------------------------------------------------------------------------- before the path
def test_butter_filters_gpu(torch_cuda, golden, capsys):
    from tests import _iir_checks as ic
    ic.check_butter_golden(golden, capsys)
    ic.check_picker_bandpass_golden(golden)
    ic.check_filtfilt_vs_oracle(n=20000)


def test_filtfilt_long_record_properties(torch_cuda):
    """16 x 2^20 float64 records kept on the device: linearity, time reversal (away from the padded edges filtfilt
    commutes with reversing the record: |H|^2 has zero phase), and the frequency response (a 60 Hz tone leaves a
    10-100 Hz band-pass scaled by |H(60 Hz)|^2 with no phase shift)."""
    from quantum_inferno_b200 import styx_fft
    torch = torch_cuda
    n = 1 << 20
    gen = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randn((16, n), generator=gen, device="cuda", dtype=torch.float64)
    y = torch.randn((16, n), generator=gen, device="cuda", dtype=torch.float64)
    f = lambda v: styx_fft.butter_bandpass(v, FS, 10.0, 100.0, 4, 0.0)
    fx, fy = f(x), f(y)
    assert isinstance(fx, torch.Tensor) and fx.shape == x.shape
    scale = float(fx.abs().max())
    assert float((f(2.0 * x - 3.0 * y) - (2.0 * fx - 3.0 * fy)).abs().max()) < 1e-11 * scale
    mid = slice(n // 4, 3 * n // 4)
    assert float((f(x.flip(1)).flip(1) - fx)[:, mid].abs().max()) < 1e-11 * scale
    k = torch.arange(n, device="cuda", dtype=torch.float64)
    tone = torch.cos(2 * np.pi * 60.0 / FS * k)[None, :]
    from scipy import signal
    b, a = signal.butter(4, [10.0 / 400.0, 100.0 / 400.0], btype="bandpass")
    gain = float(np.abs(signal.freqz(b, a, worN=[60.0], fs=FS)[1][0]) ** 2)       # zero phase, |H|^2 amplitude
    assert 0.99 < gain < 1.0
    assert float((f(tone)[0, mid] - gain * tone[0, mid]).abs().max()) < 1e-9


# ----------------------------------------------------------------------------- synthetic inputs
def test_synth_and_decimate_gpu(torch_cuda, golden, capsys):
    from tests import _synth_checks as sc
    sc.check_synth_golden(golden, capsys)
    sc.check_decimate_golden(golden)
    sc.check_noise_generators_golden(golden)
    # bench-size tone made on the device: 2^24 samples, stays a tensor, equals the closed form at probe points
    from quantum_inferno_b200.synth import benchmark_signals as bs
    sig, t, nfft, fs, fc, df = bs.well_tempered_tone(800.0, 60.0, (1 << 24) / 800.0, 0.64, dtype="float32", device_out=True)
    assert isinstance(sig, torch_cuda.Tensor) and sig.shape == (1 << 24,) and sig.dtype == torch_cuda.float32
    k = np.array([0, 1, 12345, (1 << 23) + 7, (1 << 24) - 1])
    assert np.max(np.abs(sig[k].cpu().numpy() - np.cos(2.0 * np.pi * (fc / fs) * k))) < 2e-7


def test_stft_interior_fused_path_gpu(torch_cuda):
    """The fused mean / window gather of the interior CTAs for every hop ratio, and the fallback, against the oracle."""
    from oracle import qi_oracle as orc
    from quantum_inferno_b200 import styx_fft
    k = np.arange(100000)
    x = np.random.default_rng(0).standard_normal((3, 100000)) + 3.0 + np.cos(2 * np.pi * 60 / 800 * k)
    for seg, ov, nfft in ((1024, None, None), (256, 192, 512), (128, 96, None), (300, 100, 512), (200, 150, 256), (250, 100, 256),
                          (700, 300, 1024), (2048, None, None), (1500, 1100, 2048), (512, 128, None)):
        f0, t0, z0 = orc.stft_complex_pow2(x, FS, seg, ov, nfft, alpha=0.25)
        for dtype, tol in (("float64", 1e-12), ("float32", 3e-6)):
            f, t, z = styx_fft.stft_complex_pow2(x, FS, seg, ov, nfft, alpha=0.25, dtype=dtype)
            assert z.shape == z0.shape and rel(z, z0) < tol, (seg, ov, nfft, dtype)
    f, p = styx_fft.welch_power_pow2(x, FS, 256)
    assert rel(p, orc.welch_power_pow2(x, FS, 256)[1]) < 1e-12


@pytest.mark.parametrize("order,logn,chans", [(3, 16, 3), (6, 14, 2), (12, 13, 2), (3, 18, 1)])
def test_stx_multirate_vs_oracle(torch_cuda, order, logn, chans):
    """Multirate float32 Stockwell (decimated voices, polyphase interpolation) against the oracle and the exact method."""
    from oracle import qi_oracle as orc
    from quantum_inferno_b200 import styx_stx
    n = 1 << logn
    x = np.stack([synth(n, seed=3, chan=c) for c in range(chans)])
    f, t, c = styx_stx.stx_complex_any_scale_pow2(order, x, FS, dtype="float32", method="multirate")
    _, _, p = styx_stx.stx_complex_any_scale_pow2(order, x, FS, dtype="float32", method="multirate", outputs="power")
    _, _, ce = styx_stx.stx_complex_any_scale_pow2(order, x, FS, dtype="float32")
    assert c.shape == ce.shape and c.dtype == np.complex64
    assert np.linalg.norm(c - ce) / np.linalg.norm(ce) < 1e-5
    for ch in range(chans if logn <= 16 else 1):
        f0, t0, c0 = orc.stx_complex_any_scale_pow2(order, x[ch], FS)
        assert np.array_equal(f, f0)
        assert np.linalg.norm(c[ch] - c0) / np.linalg.norm(c0) < TOL32_L2 / 10
        assert max(np.linalg.norm(c[ch, b] - c0[b]) / np.linalg.norm(c0[b]) for b in range(len(f))) < TOL32_L2 / 5
        assert l2(p[ch], np.abs(c0) ** 2) < TOL32_L2 / 5


# ----------------------------------------------------------------------------- thinly covered rows (a5, a9, a14, errors)
def test_atoms_all_dictionaries(torch_cuda, golden):
    from tests import _extra_checks as ec
    ec.check_atoms_all_dictionaries(golden)


def test_stx_general_multipass(torch_cuda, golden):
    from tests import _extra_checks as ec
    ec.check_stx_general_multipass(golden)


def test_shannon_1d_all_attributes(torch_cuda, golden):
    from tests import _extra_checks as ec
    ec.check_shannon_1d_all_attributes(golden)


def test_reference_error_paths(torch_cuda):
    from tests import _extra_checks as ec
    ec.check_reference_error_paths()


def test_stx_band_limited_routes(torch_cuda):
    from tests import _extra_checks as ec
    ec.check_stx_band_limited_routes(18, channels=3)


def test_stx_power_entropy(torch_cuda):
    from tests import _extra_checks as ec
    ec.check_stx_power_entropy(16)
