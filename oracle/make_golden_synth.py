"""
make_golden_synth.py -- generates tests/golden/synth.npz by running the UNMODIFIED reference
(quantum_inferno.synth.benchmark_signals.well_tempered_tone / quantum_chirp and utilities.sampling.decimate_*,
imported from /root/reference).  SURVEY 8(f) rank 2.  Build container only; test infrastructure.
    python oracle/make_golden_synth.py
"""
import os
import sys

import numpy as np

REF = os.environ.get("QI_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
from quantum_inferno.synth import benchmark_signals, synthetic_signals  # noqa: E402
from quantum_inferno.utilities import sampling  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
TONES = [dict(), dict(frequency_sample_rate_hz=800.0, frequency_center_hz=61.3, time_duration_s=20.48, time_fft_s=1.28,
                      use_fft_frequency=False),
         dict(frequency_sample_rate_hz=8000.0, frequency_center_hz=440.0, time_duration_s=1.0, time_fft_s=0.1)]
CHIRPS = [dict(omega=2 * np.pi * 60 / 800, order=3), dict(omega=0.3, order=12, gamma=0.7),
          dict(omega=0.9 * np.pi, order=6, gauss=False), dict(omega=0.2, order=3, oversample_scale=4)]


def main():
    d = {}
    for i, kw in enumerate(TONES):
        sig, t, nfft, fs, fc, df = benchmark_signals.well_tempered_tone(**kw)
        d[f"tone{i}_sig"], d[f"tone{i}_t"], d[f"tone{i}_meta"] = sig, t, np.array([nfft, fs, fc, df])
    for i, kw in enumerate(CHIRPS):
        wf, support = benchmark_signals.quantum_chirp(**kw)
        d[f"chirp{i}_wf"], d[f"chirp{i}_support"] = wf, np.array(support)
    x = np.random.default_rng(0).standard_normal(5000)
    d["x"] = x
    for q in (2, 4, 10):
        d[f"dec_{q}"] = sampling.decimate_timeseries(x, q)
    coll = np.random.default_rng(1).standard_normal((3, 3000))
    d["coll"], d["coll_dec_4"] = coll, sampling.decimate_timeseries_collection(coll, 4)
    # the noisy generators with the noise switched off (normal(0, std / 2**inf) = 0): their deterministic part
    d["chirp16_default"] = synthetic_signals.chirp_noise_16bit(noise_std_loss_bits=np.inf)
    d["chirp16_fc"] = synthetic_signals.chirp_noise_16bit(2 ** 13, 800.0, np.inf, frequency_center_hz=20.0)
    d["chirp_lin"], d["chirp_lin_t"] = synthetic_signals.chirp_linear_in_noise(np.inf, 800.0, 2.0, 10.0, 100.0, 0.25, 0.5)
    np.savez_compressed(os.path.join(OUT, "synth.npz"), **d)
    print("synth.npz", os.path.getsize(os.path.join(OUT, "synth.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
