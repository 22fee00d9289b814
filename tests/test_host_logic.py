"""CPU tests of the host side: band tables (bit-exact vs reference-generated vectors), integer index logic,
windows, frame bookkeeping, the C-ABI surface, and the no-fallback rule."""
import ctypes
import os
import re

import numpy as np
import pytest

from quantum_inferno_b200 import _lib, _plan, _runtime, scales_dyadic as sc
from quantum_inferno_b200.utilities import calculations, matrix, rescaling

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def test_band_tables_bit_exact(golden):
    g = golden("scales")
    for i, (fs, logn, order, nb) in enumerate(g["cases"]):
        f = sc.log_frequency_hz_from_fft_points(fs, 2 ** int(logn), order)
        assert len(f) == int(nb) and np.array_equal(f, g[f"f_{i}"]), (fs, logn, order)
    # the reference's own (commented-out) known-answer test, quantum_inferno/tests/test_scales_dyadic.py:8-21
    f = sc.log_frequency_hz_from_fft_points(100.0, 8192, scale_order=6.0, scale_ref_hz=1.0, scale_base=sc.Slice.G3)
    assert f[0] == 0.1778279410038923 and f[-1] == 39.810717055349706 and len(f) == 48


def test_g2_band_tables_bit_exact(golden):
    g = golden("scales")
    for i, (order, base, ref, lo, hi, fs) in enumerate(g["bcases"]):
        r = sc.band_frequency_low_high(order, base, ref, lo, hi, fs)
        for j, key in ((2, "band"), (4, "calg"), (5, "cgeo"), (6, "start"), (7, "end")):
            assert np.array_equal(r[j], g[f"b{i}_{key}"]), (i, key)


def test_survey_band_counts():
    # B at fs = 800 Hz
    for order, logn, nb in [(3, 16, 36), (3, 18, 42), (3, 20, 48), (3, 24, 60), (6, 22, 102), (6, 18, 78),
                            (12, 28, 263), (12, 18, 143)]:
        assert len(sc.log_frequency_hz_from_fft_points(800.0, 2 ** logn, order)) == nb


def test_scale_helpers():
    assert sc.cycles_from_order(3) == 0.75 * np.pi * 3
    assert sc.scale_order_check(-6.0) == 6.0 and sc.scale_order_check(0.1, show_warning=False) == 0.75
    assert sc.order_from_cycles(0.5) == sc.scale_order_check(1.0 / sc.M_OVER_N, show_warning=False)
    s, w = sc.scale_from_frequency_hz(3, np.array([10.0, 100.0]), 800.0)
    assert np.array_equal(w, 2.0 * np.pi * np.array([10.0, 100.0]) / 800.0)
    assert np.array_equal(s, sc.cycles_from_order(3) / w)
    assert sc.get_epsilon() == np.finfo(np.float64).eps and sc.Slice.G3 == 10.0 ** 0.3


@pytest.mark.parametrize("n", [8, 64, 256, 1000, 4096])
def test_nearest_fft_bin_matches_argmin(n):
    rng = np.random.default_rng(n)
    d = 1 / 800.0
    ff = np.fft.fftfreq(n, d)
    probes = list(rng.uniform(-500, 500, 200)) + list(ff[:8]) + list((ff[:-1] + ff[1:])[:40] / 2) + [400.0, -400.0, 1e6, 0.0]
    for f in probes:
        assert _plan.nearest_fft_bin(f, n, d) == int(np.abs(ff - f).argmin()), (n, f)


def test_stx_shift_indices_match_oracle():
    from oracle import qi_oracle as orc
    for order, n in [(3, 2048), (6, 1024), (12, 4096), (1, 256)]:
        f, bands = _plan.stx_bands(order, n, 800.0)
        assert np.array_equal(bands["shift"], orc.stx_shift_indices(order, n, 800.0))


def test_windows_and_frames(golden):
    from oracle import qi_oracle as orc
    for kind, par, n in [("tukey", 0.25, 256), ("tukey", 1.0, 1024), ("tukey", 0.0, 64), ("tukey", 0.5, 200),
                         ("gaussian", 128, 512)]:
        assert np.array_equal(_plan.periodic_window(kind, par, n), orc.window_periodic(kind, par, n))
    g = golden("stft")
    nfr, padl, ext = _plan.stft_frames(8192, 1024, 512)
    assert (nfr, padl) == (17, 512) and np.allclose(_plan.stft_time_axis(ext, 1024, 512, 800.0), g["hann_t"])
    nfr, padl, ext = _plan.stft_frames(4096, 200, 150)
    assert nfr == g["odd_z"].shape[-1]
    # configs[1]: 64 x 2^20, 1024/512 -> 2049 frames x 513 bins
    assert _plan.stft_frames(2 ** 20, 1024, 512)[0] == 2049
    with pytest.raises(ValueError):
        _plan.stft_frames(100, 16, 16)


def test_gabor_band_plan():
    f = sc.log_frequency_hz_from_fft_points(800.0, 2 ** 16, 3)
    bands, scale, omega, amp = _plan.gabor_bands(3, 2 ** 16, f, 800.0, "norm", "float64")
    assert len(bands) == 36 and bands.dtype.itemsize == 40
    assert np.array_equal(bands["omega"], 2.0 * np.pi * f / 800.0)
    assert int((bands["analytic"] == 0).sum()) == 4               # lowest ~1.2*N bands are truncated atoms
    assert np.all(np.diff(bands["analytic"]) >= 0)
    a_norm, a_spect = _plan.wavelet_amplitude(scale)
    assert np.array_equal(amp, a_norm)
    assert np.array_equal(_plan.gabor_bands(3, 2 ** 16, f, 800.0, "spect")[3], a_spect)
    assert np.all(_plan.gabor_bands(3, 2 ** 16, f, 800.0, "unit")[3] == 1.0)
    assert np.array_equal(_plan.gabor_bands(3, 2 ** 16, f, 800.0, "anything")[3], a_norm)   # silent 'norm' fallback


def test_chirp_host_helpers(golden):
    """Scalar / closed-form helpers of cwt_atoms stay on the host and are bit-exact with the reference."""
    from quantum_inferno_b200 import cwt_atoms
    g = golden("atoms")
    cases = [(3, 0, 2.0), (12, 1.0, 2.0), (6, -1.0, sc.Slice.G3), (1, 0, 2.0)]
    assert np.array_equal(np.array([cwt_atoms.chirp_mqg_from_n(*c) for c in cases]), g["mqg"])
    s, f = cwt_atoms.chirp_spectrum(np.linspace(0.0, 400.0, 64), 0.1, 3, 50.0, 800.0, 1.0)
    assert np.array_equal(s, g["spec"]) and np.array_equal(f, g["spec_f"])
    s, f = cwt_atoms.chirp_spectrum_centered(6, 80.0, 800.0, -1.0, sc.Slice.G3)
    assert np.array_equal(s, g["spec_c"]) and np.array_equal(f, g["spec_c_f"])
    m, q, gam = cwt_atoms.chirp_mqg_from_n(3)
    assert abs(m - 7.1907) < 1e-4                       # M(3) = 7.1907, not 0.75*pi*3
    assert cwt_atoms.chirp_scale(m, 10.0, 800.0) == m * 800.0 / 10.0 / (2.0 * np.pi)
    t_s, f_hz = cwt_atoms.chirp_scales_from_duration(3, 10.24)
    assert t_s == 10.24 / m and f_hz == 1 / t_s


def test_utilities():
    assert rescaling.is_power_of_two(1024) and not rescaling.is_power_of_two(1000) and not rescaling.is_power_of_two(0)
    assert rescaling.to_log2_with_epsilon(1.0) == np.log2(1.0 + np.finfo(np.float64).eps)
    assert calculations.get_num_points(800.0, 0.47, "ceil", "log2") == 9
    assert calculations.round_value(2.5) == 2 and calculations.round_value(5.0, "ceil_power_of_two") == 8
    with pytest.raises(ValueError):
        calculations.round_value(1.0, "nope")
    a = np.arange(6.0).reshape(2, 3)
    assert np.array_equal(matrix.d0tile_x_d0d1(np.array([2.0, 3.0]), a), a * np.array([[2.0], [3.0]]))
    assert np.array_equal(matrix.d1tile_x_d0d1(np.array([1.0, 2.0, 3.0]), a), a * np.array([1.0, 2.0, 3.0]))
    with pytest.raises(TypeError):
        matrix.d0tile_x_d0d1(np.array([1.0, 2.0, 3.0]), a)


def test_c_abi_exports_every_declared_symbol():
    """The CUDA library loads without a GPU and exports exactly what include/qi_b200.h declares."""
    if not os.path.exists(_lib.LIB_PATH):
        pytest.skip("libqi_b200.so not built (run __graft_entry__.build())")
    hdr = open(os.path.join(ROOT, "include", "qi_b200.h")).read()
    declared = set(re.findall(r"\b(qi_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    lib = _lib.bind(ctypes.CDLL(_lib.LIB_PATH))
    assert lib.qi_abi_version() == _lib.QI_ABI_VERSION
    assert lib.qi_error_string(-2) == b"workspace too small"
    # struct mirrors
    assert _lib.ATOM_BAND.itemsize == 40 and _lib.STX_BAND.itemsize == 16
    # argument validation happens before any CUDA call
    assert lib.qi_fft_c2c(None, None, 1, 4, 0, 0, None) == -1
    assert lib.qi_cwt_workspace_bytes(1, 4096, 27, 4, 1, 0, 1) > 4096 * 2 * 16 * 2


def test_host_side_of_the_next_rows():
    """Compiled host code and host bookkeeping of the SURVEY 8(f) rows that needs no GPU: scipy's minimum-distance
    peak selection (qi_select_peaks_by_distance), struct mirrors and argument checks of the new entry points, and the
    exact transfer-function -> sections factoring of the pre-filters."""
    if not os.path.exists(_lib.LIB_PATH):
        pytest.skip("libqi_b200.so not built (run __graft_entry__.build())")
    from oracle import qi_oracle as orc
    from quantum_inferno_b200 import _iir
    lib = _lib.bind(ctypes.CDLL(_lib.LIB_PATH))
    rng = np.random.default_rng(4)
    for n, dist in ((0, 3), (1, 3), (500, 1), (500, 7), (500, 40), (2000, 1000)):
        peaks = np.sort(rng.choice(20000, size=n, replace=False)).astype(np.int64)
        prio = rng.standard_normal(n)
        order = np.ascontiguousarray(np.argsort(prio), dtype=np.int64)
        keep = np.zeros(n, dtype=np.uint8)
        assert lib.qi_select_peaks_by_distance(peaks.ctypes.data, order.ctypes.data, n, dist, keep.ctypes.data) == 0
        assert np.array_equal(keep.astype(bool), orc.select_by_peak_distance(peaks, prio, dist)), (n, dist)
    assert lib.qi_select_peaks_by_distance(None, None, 5, 3, None) == -1
    assert _lib.IIR_FILTER.itemsize == 8 + 17 * 8 * 2 + 48 * 8 + 16 * 8
    assert lib.qi_filtfilt_workspace_bytes(1, 4096, 27, 8) > 8 * (4096 + 54)
    assert lib.qi_filtfilt_workspace_bytes(1, 4096, 27, 17) == 0
    assert lib.qi_subsample(None, 1, 8, 8, 2, 1, 0, None, 4, None) == -1
    assert lib.qi_synth_chirp(1, 8, 8, 0, 0.0, 0.1, 0.0, 0.0, 0, 1, None, None, None) == -1
    # multirate Stockwell planning (host side of qi_stx_multirate): workspace grows with the record, refuses short records
    from quantum_inferno_b200 import _plan
    f, sb = _plan.stx_bands(3, 1 << 16, 800.0)
    sb = np.ascontiguousarray(sb, dtype=_lib.STX_BAND)
    w16 = lib.qi_stx_multirate_workspace_bytes(2, 1 << 16, sb.ctypes.data, len(sb))
    assert w16 > 2 * (1 << 16) * 8 * 2 and lib.qi_stx_multirate_workspace_bytes(2, 2048, sb.ctypes.data, len(sb)) == 0
    assert lib.qi_stx_multirate(None, 2, 1 << 16, 1 << 16, sb.ctypes.data, len(sb), None, None, None, 0, None) == -1
    # factoring the reference's rounded taps: sections multiply back to the taps (float64 product of the sections is
    # good to 1e-9 for these designs; the exactness itself is asserted against the long-double recursion elsewhere)
    from scipy import signal
    for b, a in (signal.butter(4, [0.025, 0.25], btype="bandpass"), signal.butter(5, 0.3), signal.butter(2, 0.1, btype="highpass")):
        sos = _iir.tf2sos_exact(b, a)
        bb, aa = np.ones(1), np.ones(1)
        for sec in sos:
            bb, aa = np.convolve(bb, sec[:3]), np.convolve(aa, sec[3:])
        assert np.allclose(np.trim_zeros(bb, "b"), b, rtol=1e-7, atol=1e-12 * np.max(np.abs(b)))
        assert np.allclose(np.trim_zeros(aa, "b"), a, rtol=1e-7, atol=1e-9)
        assert np.all(sos[:, 3] == 1.0) and sos.shape == ((len(a)) // 2, 6)


def test_no_cpu_fallback():
    """Without a CUDA device the product path raises; it never routes through the oracle or numpy."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from quantum_inferno_b200 import styx_cwt
    _runtime._runtime = None
    with pytest.raises(RuntimeError):
        styx_cwt.cwt_complex_any_scale_pow2(3, np.zeros(256), 800.0)
    pkg = os.path.join(ROOT, "quantum_inferno_b200")
    src = "".join(open(os.path.join(d, fn)).read() for d in (pkg, os.path.join(pkg, "utilities"), os.path.join(pkg, "synth"))
                  for fn in os.listdir(d) if fn.endswith(".py"))
    assert "qi_oracle" not in src and "libqi_emul" not in src and "import oracle" not in src


def test_band_shard_cost_balance():
    """distributed.band_shard: contiguous cover of the band table, at least one band per rank, balanced on the per-level
    cost model of the multirate path (a level-0 band costs several deep bands)."""
    from quantum_inferno_b200 import distributed
    n, order = 1 << 24, 12
    f = sc.log_frequency_hz_from_fft_points(800.0, n, order)
    cost = distributed.band_cost(order, n, f, 800.0, {"dtype": "float32"})
    assert len(cost) == len(f) and min(cost) >= 1.0 and cost[-1] > cost[0]          # high bands (level 0) cost more
    assert distributed.band_cost(order, n, f, 800.0, {"dtype": "float64"}) is None     # exact method: equal cost
    for world in (2, 3, 8):
        edges = [distributed.band_shard(len(f), r, world, cost) for r in range(world)]
        assert edges[0][0] == 0 and edges[-1][1] == len(f)
        assert all(a[1] == b[0] for a, b in zip(edges[:-1], edges[1:])) and all(b1 > b0 for b0, b1 in edges)
        loads = [sum(cost[b0:b1]) for b0, b1 in edges]
        assert max(loads) < 1.25 * sum(cost) / world
        counts = [b1 - b0 for b0, b1 in edges]
        assert counts[0] > counts[-1]                                               # fewer of the expensive bands
    with pytest.raises(ValueError):
        distributed.band_shard(3, 0, 4)
    assert [distributed.channel_shard(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]


def test_multirate_level_assignment_keeps_the_decimator_error_at_its_ripple():
    """_plan.MR_KAPPA: a band may sit at level l if the LAST half-band decimation stage (7-tap minimax filter of
    csrc/qi_halfband_coeffs.h) neither droops nor aliases where the band still answers -- the weighted error
    max |H(theta) - 1| exp(-((theta - omega) s)^2 / 2) must stay at the filter's own pass-band ripple."""
    import os
    import re
    from quantum_inferno_b200 import _plan, scales_dyadic as scales
    hdr = open(os.path.join(os.path.dirname(_plan.__file__), "csrc", "qi_halfband_coeffs.h")).read()
    taps = np.array([float(v) for v in re.search(r"qi_hb_taps\[QI_HB_CLASSES\]\[QI_HB_MAX_TAPS\] = \{\s*\{([^}]*)\}", hdr).group(1).split(",")])
    th = np.linspace(0.0, np.pi, 4097)
    h_err = np.abs(0.5 + 2.0 * sum(c * np.cos((2 * t + 1) * th) for t, c in enumerate(taps)) - 1.0)
    ripple = h_err[th <= np.pi / 4].max()
    assert ripple < 1.1e-6
    for order, log2n in ((1.5, 20), (3, 24), (6, 22), (12, 24), (24, 20)):
        n = 1 << log2n
        f = scales.log_frequency_hz_from_fft_points(800.0, n, order)
        bands, scale, omega, _ = _plan.multirate_bands(order, n, f, 800.0)
        worst = 0.0
        for lv, s, w in zip(bands["level"], scale, omega):
            if lv == 0:
                continue
            up = 2.0 ** (int(lv) - 1)                      # the band in the units of the level the last stage reads
            worst = max(worst, float((h_err * np.exp(-0.5 * ((th - w * up) * (s / up)) ** 2)).max()))
            assert (w + 5.2 / s) * 2.0 ** int(lv) < np.pi  # and its kernel is representable at its own level's rate
        assert worst < 1.5 * ripple, (order, worst, ripple)
