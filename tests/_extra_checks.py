"""Shared assertions for the thinly covered rows of SURVEY 8(a) -- a5 atoms, a9 general Stockwell transform, a14 1-D
Shannon classes, and the reference's ValueError paths -- run through the CPU emulator (tests/test_emul_api.py) and on the
B200 (tests/test_gpu_parity.py) against tests/golden/extra.npz (oracle/make_golden_extra.py)."""
import numpy as np
import pytest

FS = 800.0
ATOM_CASES = [(3, 512, "norm"), (3, 1000, "spect"), (6, 2048, "unit"), (12, 1000, "norm"), (6, 512, "spect"), (3, 2048, "unit")]
ATOM_FREQS = np.array([2.5, 40.0, 310.0])
INFO_LENGTHS = [1000, 777, 1 << 13]


def rel(a, b):
    return float(np.max(np.abs(np.asarray(a) - b)) / np.max(np.abs(b)))


def check_atoms_all_dictionaries(golden):
    """a5: styx_cwt.wavelet_centered_4cwt (reference styx_cwt.py:113-144), three dictionaries x three sizes x three orders."""
    from quantum_inferno_b200 import styx_cwt
    g = golden("extra")
    for i, (order, n, dic) in enumerate(ATOM_CASES):
        atoms, t_s, scale, omega, amp = styx_cwt.wavelet_centered_4cwt(order, n, ATOM_FREQS, FS, dic)
        assert atoms.shape == (3, n) and atoms.dtype == np.complex128
        assert rel(atoms, g[f"atoms{i}"]) < 1e-12, (order, n, dic)
        assert np.array_equal(t_s, g[f"atoms{i}_t"])
        assert np.array_equal(scale[:, 0], g[f"atoms{i}_scale"]) and np.array_equal(omega[:, 0], g[f"atoms{i}_omega"])
        assert np.array_equal(amp[:, 0], g[f"atoms{i}_amp"])
        a32 = styx_cwt.wavelet_centered_4cwt(order, n, ATOM_FREQS, FS, dic, dtype="float32")[0]
        assert a32.dtype == np.complex64 and rel(a32, g[f"atoms{i}"]) < 2e-6


def check_stx_general_multipass(golden):
    """a9: styx_stx.tfr_stx_fft (reference styx_stx.py:52-192) at 2048 samples = two FFT passes, and its error paths."""
    from quantum_inferno_b200 import styx_stx
    g = golden("extra")
    x = g["stx_x"]
    for tag, kw in [("lin", dict(frequency_min=10.0, frequency_max=100.0, frequency_step=5.0)),
                    ("geo", dict(frequency_min=10.0, frequency_max=100.0, is_geometric=True, scale_order_input=3.0))]:
        tfr, psd, f, f_fft, win = styx_stx.tfr_stx_fft(x, 1 / FS, n_fft_in=2048, **kw)
        assert np.array_equal(f, g[f"stx_{tag}_f"]) and np.array_equal(f_fft, g[f"stx_{tag}_ffft"])
        assert tfr.shape == g[f"stx_{tag}_tfr"].shape and rel(tfr, g[f"stx_{tag}_tfr"]) < 1e-10, tag
        assert rel(psd, g[f"stx_{tag}_psd"]) < 1e-10 and rel(win[::4], g[f"stx_{tag}_win0"]) < 1e-12
    # reference styx_stx.py:30-31: a transform shorter than the record is an error
    with pytest.raises(ValueError):
        styx_stx.sig_pad_up_to_pow2(x, 1024)
    with pytest.raises(ValueError):
        styx_stx.tfr_stx_fft(x, 1 / FS, n_fft_in=1024)


def check_shannon_1d_all_attributes(golden):
    """a14: ShannonTDR / ShannonFFT (reference tfr_info.py:106-200): every attribute, record lengths that are not powers
    of two included (scipy's rfft takes any length)."""
    from quantum_inferno_b200 import tfr_info
    g = golden("extra")
    for n in INFO_LENGTHS:
        tdr, ff = tfr_info.shannon_tdr_fft(g[f"info{n}_x"])
        for tag, obj in (("tdr", tdr), ("fft", ff)):
            assert rel(obj.sig, g[f"info{n}_{tag}_sig"]) < 1e-10, (n, tag)
            assert rel(obj.marginal, g[f"info{n}_{tag}_marginal"]) < 1e-10
            # information of near-empty cells: log2 of (marginal + eps32) with marginal ~ eps64 * peak
            assert np.max(np.abs(obj.info - g[f"info{n}_{tag}_info"])) < 1e-9, (n, tag)
            assert rel(obj.entropy, g[f"info{n}_{tag}_entropy"]) < 1e-10
            assert np.max(np.abs(obj.isnr - g[f"info{n}_{tag}_isnr"])) < 1e-9
            assert rel(obj.esnr, g[f"info{n}_{tag}_esnr"]) < 1e-10
            assert abs(obj.ref_entropy - float(g[f"info{n}_{tag}_ref"])) < 1e-15
        assert np.array_equal(ff.frequency, g[f"info{n}_fft_freq"])
        # the unwrapped phase is only defined where the spectrum is not numerically zero
        strong = np.abs(g[f"info{n}_fft_sig"]) > 1e-6 * np.abs(g[f"info{n}_fft_sig"]).max()
        d = (ff.angle_rads - g[f"info{n}_fft_angle"])[strong]
        assert np.max(np.abs(d - 2 * np.pi * np.round(d / (2 * np.pi)))) < 1e-8


def check_reference_error_paths():
    """ValueError conventions of the reference on the live runtime: styx_fft.py:42-45 (record shorter than the FFT window),
    cwt_atoms.py:436-437 (unknown cwt_type)."""
    from quantum_inferno_b200 import cwt_atoms, styx_fft
    x = np.cos(2 * np.pi * 60.0 / FS * np.arange(256))
    with pytest.raises(ValueError):
        styx_fft.stft_from_sig(x, FS, 3, center_frequency_hz=0.5)       # needs a window far longer than 256 samples
    with pytest.raises(ValueError):
        cwt_atoms.cwt_chirp_complex(3, x, 10.0, FS, cwt_type="nope")


def check_stx_band_limited_routes(log2n, channels=1):
    """styx_stx.stx_complex_any_scale_pow2 (reference styx_stx.py:195-236) on records long enough for the band-limited
    route of csrc/qi_stx.cu (decimated voices + Kaiser interpolation): against the full-length passes and the oracle."""
    from oracle import qi_oracle as orc
    from quantum_inferno_b200 import styx_stx
    n = 1 << log2n
    k = np.arange(n)
    x = np.stack([np.cos(2 * np.pi * (40.0 + 7 * c) / FS * k) + 0.3 * np.random.default_rng(c).standard_normal(n)
                  for c in range(channels)])
    _, _, ref = orc.stx_complex_any_scale_pow2(3, x[-1], FS)
    for dtype, tol in (("float64", 1e-10), ("float32", 2e-5)):
        f, _, c = styx_stx.stx_complex_any_scale_pow2(3, x, FS, dtype=dtype)
        _, _, cp = styx_stx.stx_complex_any_scale_pow2(3, x, FS, dtype=dtype, method="plain")
        per_band = np.max(np.abs(c - cp), axis=-1) / np.max(np.abs(cp), axis=-1)
        assert per_band.max() < tol, (dtype, per_band)
        assert np.any(per_band > 0)                               # some bands did take another route
        err = np.max(np.abs(c[-1] - ref), axis=-1) / np.max(np.abs(ref), axis=-1)
        assert err.max() < tol, (dtype, err)
        _, _, p = styx_stx.stx_complex_any_scale_pow2(3, x, FS, dtype=dtype, outputs="power")
        assert np.max(np.abs(p[-1] - np.abs(ref) ** 2)) / np.max(np.abs(ref) ** 2) < 4 * tol


def check_stx_power_entropy(log2n):
    """cwt_entropy.stx_power_entropy: Stockwell power / information / entropy without the complex plane, against the
    reference composition (styx_stx.py:195-236 + tfr_info.py:203-236) restated by the oracle."""
    from oracle import qi_oracle as orc
    from quantum_inferno_b200 import cwt_entropy
    n = 1 << log2n
    k = np.arange(n)
    x = np.cos(2 * np.pi * 55.0 / FS * k) + 0.3 * np.random.default_rng(11).standard_normal(n)
    _, _, ref = orc.stx_complex_any_scale_pow2(3, x, FS)
    p_ref = np.abs(ref) ** 2
    s_ref = p_ref.sum()
    pdf = p_ref / s_ref
    info_ref = -np.log2(pdf + np.finfo(np.float64).eps)
    ent_ref = float(np.sum(pdf * info_ref))
    for dtype, tol_p, tol_bits in (("float64", 1e-10, 1e-9), ("float32", 1e-4, 1e-3)):
        r = cwt_entropy.stx_power_entropy(3, x, FS, dtype=dtype)
        p = np.asarray(r.power[0].cpu() if hasattr(r.power, "cpu") else r.power[0], dtype=np.float64)
        assert np.linalg.norm(p - p_ref) / np.linalg.norm(p_ref) < tol_p
        tot = float(r.total_power[0])
        assert abs(tot - s_ref) / s_ref < tol_p
        ent = float(r.entropy_bits()[0])
        assert abs(ent - ent_ref) < tol_bits, (dtype, ent, ent_ref)
        info = np.asarray(r.info[0].cpu() if hasattr(r.info, "cpu") else r.info[0], dtype=np.float64)
        strong = p_ref > 1e-3 * p_ref.max()
        assert np.max(np.abs(info - info_ref)[strong]) < (1e-8 if dtype == "float64" else 1e-3)
