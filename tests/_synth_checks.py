"""Shared assertions for the device-side synthetic inputs and decimation, run through the CPU
emulator (tests/test_emul_api.py) and on the B200 (tests/test_gpu_parity.py)."""
import numpy as np
import pytest

TONES = [dict(), dict(frequency_sample_rate_hz=800.0, frequency_center_hz=61.3, time_duration_s=20.48, time_fft_s=1.28,
                      use_fft_frequency=False),
         dict(frequency_sample_rate_hz=8000.0, frequency_center_hz=440.0, time_duration_s=1.0, time_fft_s=0.1)]
CHIRPS = [dict(omega=2 * np.pi * 60 / 800, order=3), dict(omega=0.3, order=12, gamma=0.7),
          dict(omega=0.9 * np.pi, order=6, gauss=False), dict(omega=0.2, order=3, oversample_scale=4)]


def peak_rel(a, b):
    return float(np.max(np.abs(np.asarray(a) - b)) / np.max(np.abs(b)))


def check_synth_golden(golden, capsys):
    from quantum_inferno_b200.synth import benchmark_signals as bs
    g = golden("synth")
    for i, kw in enumerate(TONES):
        sig, t, nfft, fs, fc, df = bs.well_tempered_tone(**kw)
        # the phase (2 pi f_c) * n is formed exactly as numpy does; cos may differ by an ulp of the libm in use
        assert sig.dtype == np.float64 and np.max(np.abs(sig - g[f"tone{i}_sig"])) < 1e-11, i   # |arg| up to 1e5 rad
        assert np.array_equal(t, g[f"tone{i}_t"]) and np.array_equal([nfft, fs, fc, df], g[f"tone{i}_meta"])
    assert "power of two" in capsys.readouterr().out                           # tone 2 warns like the reference
    s32 = bs.well_tempered_tone(dtype="float32")[0]
    assert s32.dtype == np.float32 and np.max(np.abs(s32 - g["tone0_sig"])) < 1e-7
    with pytest.raises(NotImplementedError):
        bs.well_tempered_tone(add_noise_taper_aa=True)
    for i, kw in enumerate(CHIRPS):
        wf, support = bs.quantum_chirp(**kw)
        assert wf.dtype == np.complex128 and support == int(g[f"chirp{i}_support"])
        assert wf.shape == g[f"chirp{i}_wf"].shape and peak_rel(wf, g[f"chirp{i}_wf"]) < 1e-12, i


def check_decimate_golden(golden):
    from quantum_inferno_b200.utilities import sampling
    g = golden("synth")
    for q in (2, 4, 10):
        y = sampling.decimate_timeseries(g["x"], q)
        assert y.shape == g[f"dec_{q}"].shape and peak_rel(y, g[f"dec_{q}"]) < 1e-11, q
    y = sampling.decimate_timeseries_collection(g["coll"], 4)
    assert y.shape == g["coll_dec_4"].shape and peak_rel(y, g["coll_dec_4"]) < 1e-11
    with pytest.raises(ValueError):
        sampling.decimate_timeseries(g["x"][:27], 2)                           # the reference needs 28 samples or more


def check_noise_generators_golden(golden):
    """chirp_noise_16bit / chirp_linear_in_noise / white_noise_fbits (reference synth/synthetic_signals.py:53-81, :127-169):
    the deterministic part against the reference with the noise switched off, the noise by its seed and its statistics."""
    from quantum_inferno_b200.synth import synthetic_signals as ss
    g = golden("synth")
    y = ss.chirp_noise_16bit(noise_std_loss_bits=np.inf)
    assert y.dtype == np.float16 and y.shape == g["chirp16_default"].shape
    # float16 output: one half-precision ulp where a float64 value sits on a rounding boundary
    assert np.max(np.abs(y.astype(np.float64) - g["chirp16_default"].astype(np.float64))) <= 2.0 ** -10
    assert np.mean(y == g["chirp16_default"]) > 0.999
    y = ss.chirp_noise_16bit(2 ** 13, 800.0, np.inf, frequency_center_hz=20.0)
    assert np.max(np.abs(y.astype(np.float64) - g["chirp16_fc"].astype(np.float64))) <= 2.0 ** -10
    w, t = ss.chirp_linear_in_noise(np.inf, 800.0, 2.0, 10.0, 100.0, 0.25, 0.5)
    assert w.dtype == np.float64 and np.array_equal(t, g["chirp_lin_t"])
    assert np.max(np.abs(w - g["chirp_lin"])) < 1e-11
    # seeded noise: repeatable, of the requested level, uncorrelated with the sweep
    a, _ = ss.chirp_linear_in_noise(3.0, 800.0, 2.0, 10.0, 100.0, 0.25, 0.5, seed=7)
    b, _ = ss.chirp_linear_in_noise(3.0, 800.0, 2.0, 10.0, 100.0, 0.25, 0.5, seed=7)
    c, _ = ss.chirp_linear_in_noise(3.0, 800.0, 2.0, 10.0, 100.0, 0.25, 0.5, seed=8)
    assert np.array_equal(a, b) and not np.array_equal(a, c)
    noise = a - g["chirp_lin"]
    want = np.std(g["chirp_lin"]) / 2.0 ** 3
    assert abs(np.std(noise) / want - 1.0) < 0.05 and abs(np.mean(noise)) < 4 * want / np.sqrt(noise.size)
    assert abs(np.corrcoef(noise, g["chirp_lin"])[0, 1]) < 0.1
    n = ss.white_noise_fbits(g["x"], 2.0, seed=1)
    assert n.shape == (g["x"].size,) and abs(np.std(n) / (np.std(g["x"]) / 4.0) - 1.0) < 0.05
    assert np.array_equal(n, ss.white_noise_fbits(g["x"], 2.0, seed=1))
