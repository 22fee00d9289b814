"""
Secondary measurements for the other BASELINE.json configs (the headline line is bench.py).  One GPU, the per-GPU
share of each config, CUDA-event timing, inputs resident in HBM, 3 warm-ups.  Prints one JSON object per config, with
the SM clocks / throttle reasons nvidia-smi reported while it ran (``clocks``, as in bench.py).

    python tools/bench_configs.py [cfg1 cfg2 cfg3 cfg4 cfg5]
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from bench import FS, ClockSampler, synth_batch_torch  # noqa: E402
from quantum_inferno_b200 import cwt_entropy, styx_cwt, styx_fft, styx_stx  # noqa: E402

DEV = torch.device("cuda", 0)


def spin(ms=300.0):
    """Keep the GPU busy for a while so that the short measurements below run at the loaded clock."""
    a = torch.randn(4096, 4096, device=DEV)
    t0 = time.time()
    while (time.time() - t0) * 1e3 < ms:
        for _ in range(10):
            a = (a @ a).tanh_()
        torch.cuda.synchronize()


def timed(fn, reps=10, warm=3):
    """CUDA-event time per call: best of three passes of `reps` calls (the first pass after a cold start is host-bound:
    allocator growth, lazy module loads), inputs resident, clocks warmed."""
    for _ in range(warm):
        fn()
    spin()
    best = None
    for _ in range(3):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            out = fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        best = ms if best is None else min(best, ms)
    return best, out


def cfg1():
    """single channel 2^16 @ 800 Hz, order 3 CWT + power/entropy: the drop-in fp64 API and the fused fp32 path."""
    x = synth_batch_torch(torch, 1 << 16, [0], DEV)[0]
    x64 = x.double()
    ms64, (f, t, c) = timed(lambda: styx_cwt.cwt_complex_any_scale_pow2(3, x64, FS))
    ms32, r = timed(lambda: cwt_entropy.cwt_power_entropy(3, x, FS, dtype="float32"))
    cells = c.shape[-2] * c.shape[-1]
    return {"config": "cfg1: 1 x 2^16, N=3 CWT", "bands": int(c.shape[-2]),
            "fp64_complex_tfr_ms": ms64, "fp64_cells_per_s": cells / ms64 * 1e3,
            "fp32_fused_power_info_entropy_ms": ms32, "fp32_cells_per_s": cells / ms32 * 1e3}


def cfg2():
    """STFT Hann 1024 / 50 % on 64 x 2^20 (fp32 and fp64)."""
    x = synth_batch_torch(torch, 1 << 20, list(range(64)), DEV)
    out = {"config": "cfg2: STFT Hann 1024/512, 64 x 2^20"}
    for dt, xx in (("float32", x), ("float64", x.double())):
        ms, (f, t, z) = timed(lambda: styx_fft.stft_complex_pow2(xx, FS, 1024, alpha=1.0, dtype=dt))
        cells = z.shape[0] * z.shape[1] * z.shape[2]
        bytes_alg = cells * (8 if dt == "float32" else 16) + xx.numel() * xx.element_size()
        out[dt] = {"ms": ms, "cells_per_s": cells / ms * 1e3, "samples_per_s": x.numel() / ms * 1e3,
                   "alg_GBps": bytes_alg / ms / 1e6, "shape": list(z.shape)}
    return out


def cfg3():
    """Stockwell on 16 x 2^18, orders 3/6/12, fp32 and fp64."""
    x = synth_batch_torch(torch, 1 << 18, list(range(16)), DEV)
    out = {"config": "cfg3: STX 16 x 2^18"}
    for order in (3, 6, 12):
        for dt, xx in (("float32", x), ("float64", x.double())):
            if order == 12 and dt == "float64":
                xx = xx[:8]                                   # 143 bands x 2^18 complex128 x 16 ch = 9.6 GB + workspace
            ms, (f, t, z) = timed(lambda: styx_stx.stx_complex_any_scale_pow2(order, xx, FS, dtype=dt), reps=3)
            cells = z.shape[0] * z.shape[1] * z.shape[2]
            out[f"order{order}_{dt}"] = {"ms": ms, "cells_per_s": cells / ms * 1e3, "bands": int(z.shape[1]),
                                          "channels": int(z.shape[0])}
            if dt == "float32":       # multirate opt-in: decimated voices + polyphase interpolation (complex64 out: 8 B/cell)
                ze = z
                ms, (f, t, z) = timed(lambda: styx_stx.stx_complex_any_scale_pow2(order, xx, FS, dtype=dt, method="multirate"), reps=3)
                err = float(torch.linalg.vector_norm(z - ze) / torch.linalg.vector_norm(ze))
                out[f"order{order}_float32_multirate"] = {"ms": ms, "cells_per_s": cells / ms * 1e3, "alg_GBps": cells * 8 / ms / 1e6,
                                                          "rel_l2_vs_exact": err}
                del ze
            del z
            torch.cuda.empty_cache()
    return out


def cfg4():
    """order 6 CWT + info/entropy, 256 ch x 2^22 over 8 GPUs -> the per-GPU share: 32 ch x 2^22 (102 bands), in 4 groups."""
    n, ch = 1 << 22, 8
    x = synth_batch_torch(torch, n, list(range(ch)), DEV)
    ms, r = timed(lambda: cwt_entropy.cwt_power_entropy(6, x, FS, dtype="float32"), reps=3)
    cells = ch * r.power.shape[1] * n
    return {"config": "cfg4 share: N=6 CWT + info/entropy, 8 of the 32 ch/GPU x 2^22 per call", "bands": int(r.power.shape[1]),
            "ms_per_8ch": ms, "cells_per_s": cells / ms * 1e3, "ms_per_gpu_share_32ch": 4 * ms}


def cfg5():
    """N=12 CWT of one 2^28 record, band-sharded over 8 GPUs -> this GPU's 33-band share (planes 8 x 2^28 x 33 B)."""
    n = 1 << 28
    x = synth_batch_torch(torch, n, [0], DEV)[0]
    nb = 263
    sl = (0, 33)
    tot = torch.zeros(1, dtype=torch.float64, device=DEV)
    ms, r = timed(lambda: cwt_entropy.cwt_power_entropy(12, x, FS, dtype="float32", band_slice=sl,
                                                        allreduce=lambda t: t), reps=2, warm=1)
    cells = (sl[1] - sl[0]) * n
    return {"config": "cfg5 share: N=12 CWT, 1 x 2^28, bands 0..32 of 263 (lowest = deepest levels)", "ms": ms,
            "cells_per_s": cells / ms * 1e3}


def cfg6():
    """utilities/short_time_fft on the config-2 shape (64 x 2^20, Tukey 1024/512): |STFT| of the
    detrended slices, and the complex STFT -> inverse STFT round trip."""
    from quantum_inferno_b200.utilities import short_time_fft as stf
    x = synth_batch_torch(torch, 1 << 20, list(range(64)), DEV)
    out = {"config": "cfg6: stft_tukey / istft_tukey 1024/512, 64 x 2^20"}
    for dt, xx in (("float32", x), ("float64", x.double())):
        es = 4 if dt == "float32" else 8
        ms, (f, t, mag) = timed(lambda: stf.stft_tukey(xx, FS, 0.25, 1024, 512, dtype=dt))
        cells = mag.numel()
        obj = stf.get_stft_object_tukey(FS, 0.25, 1024, 512, dtype=dt)
        spec = obj.stft(xx)
        ms_i, (ts, xr) = timed(lambda: stf.istft_tukey(spec, FS, 0.25, 1024, 512, dtype=dt))
        err = float((xr - xx[:, :xr.shape[-1]]).abs().max())
        out[dt] = {"stft_tukey_ms": ms, "cells_per_s": cells / ms * 1e3,
                   "stft_alg_GBps": (cells * es + xx.numel() * es) / ms / 1e6,
                   "istft_ms": ms_i, "istft_alg_GBps": (spec.numel() * 2 * es + xr.numel() * es) / ms_i / 1e6,
                   "round_trip_max_abs_err": err}
    return out


def cfg7():
    """the display / picking summaries of a TFR that stays in HBM.  subsample_2d of one channel's
    [60, 2^24] fp32 power plane (4.0 GB, larger than L2) to the reference's 2^19-sample mesh (factor 32) and to a
    coarse overview (factor 2048), and find_peaks_with_bits on a 2^26-sample fp32 record.  Algorithmic bytes: one read
    of the input + the reduced output."""
    from quantum_inferno_b200.utilities import picker, sampling
    gen = torch.Generator(device=DEV).manual_seed(3)
    p = torch.rand((60, 1 << 24), generator=gen, device=DEV, dtype=torch.float32) ** 4
    out = {"config": "cfg7: subsample_2d [60, 2^24] fp32; find_peaks_with_bits 2^26 fp32", "hbm_peak_GBps": 6547.2}
    for f in (32, 2048):
        for m in ("average", "max", "median", "nth"):
            ms, y = timed(lambda: sampling.subsample_2d(p, f, m))
            nbytes = (p.numel() if m != "nth" else y.numel() * 8) * 4 + y.numel() * 4   # nth touches one 32 B sector per output
            out[f"{m}_x{f}"] = {"ms": ms, "alg_GBps": nbytes / ms / 1e6, "samples_per_s": p.numel() / ms * 1e3}
    del p
    k = torch.arange(1 << 26, device=DEV, dtype=torch.float32)
    x = torch.randn(1 << 26, generator=gen, device=DEV) * 0.05
    for c in range(1, 64):
        x += (1.0 + 0.5 * (c % 3)) * torch.exp(-0.5 * ((k - c * (1 << 20)) / 2000.0) ** 2)
    ms, pk = timed(lambda: picker.find_peaks_with_bits(x, FS, "amplitude", 1, 0.1))
    out["find_peaks_with_bits"] = {"ms": ms, "peaks": int(len(pk)), "samples_per_s": x.numel() / ms * 1e3,
                                   "alg_GBps": x.numel() * 4 * 4 / ms / 1e6,
                                   "note": "four streaming passes over the record: log2 (read+write), extrema, local maxima"}
    return out


def cfg8():
    """the Butterworth pre-filters in front of the path.  butter_bandpass (order 4 -> 4 sections,
    Tukey taper fused) on 16 x 2^22 records, float64 and float32 I/O (fp64 arithmetic), and the picker's order-7
    sosfiltfilt on one 2^24 record.  Algorithmic bytes: record read + result written (2 x element size per sample);
    real traffic is ~48 B per sample (three kernels per direction, fp64 intermediate)."""
    from quantum_inferno_b200 import styx_fft
    from quantum_inferno_b200.utilities import picker
    x = synth_batch_torch(torch, 1 << 22, list(range(16)), DEV)
    out = {"config": "cfg8: butter_bandpass 10-100 Hz order 4, 16 x 2^22; picker.apply_bandpass order 7, 1 x 2^24"}
    for dt, xx in (("float32", x), ("float64", x.double())):
        ms, y = timed(lambda: styx_fft.butter_bandpass(xx, FS, 10.0, 100.0))
        out[dt] = {"ms": ms, "samples_per_s": xx.numel() / ms * 1e3, "alg_GBps": 2 * xx.numel() * xx.element_size() / ms / 1e6,
                   "traffic_model_GBps": 48.0 * xx.numel() / ms / 1e6}
    del x, xx
    r = synth_batch_torch(torch, 1 << 24, [0], DEV)[0].double()
    ms, y = timed(lambda: picker.apply_bandpass(r, (100.0, 200.0), FS, 7))
    out["sos_order7"] = {"ms": ms, "samples_per_s": r.numel() / ms * 1e3, "alg_GBps": 16.0 * r.numel() / ms / 1e6}
    return out


if __name__ == "__main__":
    which = sys.argv[1:] or ["cfg1", "cfg2", "cfg3", "cfg4"]
    for name in which:
        t0 = time.time()
        sampler = ClockSampler(0)                 # SM clock and throttle reasons while this configuration runs
        sampler.start()
        res = globals()[name]()
        res["clocks"] = sampler.stop()
        res["wall_s"] = time.time() - t0
        print(json.dumps(res), flush=True)
        torch.cuda.empty_cache()
