"""
Fused order-N Gabor CWT + power + Shannon information / entropy -- the composition the reference spells as

    f, t, cwt = styx_cwt.cwt_complex_any_scale_pow2(N, sig, fs)          # styx_cwt.py:147-198
    power     = np.abs(cwt) ** 2
    shannon   = tfr_info.shannon_stft_from_tfr_power(power)              # tfr_info.py:231-236

kept on the device end to end: the complex TFR is never materialised (at the north-star size it would be
~1 TB), the power plane is written once by the last inverse-FFT pass together with the fp64 per-band sums,
and one streaming pass turns it into the information plane -log2(P/S + eps64) and the per-band entropy sums.

Multi-GPU: records (channels) shard with no communication; one long record can instead be sharded by band,
in which case the only exchange is one all-reduce of the per-record total power S (``allreduce=``).
"""
from dataclasses import dataclass
from typing import Callable, Optional

import numpy as np

from . import _driver, _plan
from . import scales_dyadic as scales
from ._runtime import dtype_name, get_runtime


@dataclass
class CwtEntropy:
    """Device-resident result.  Shapes: C records, B (local) bands, N samples."""
    frequency_hz: np.ndarray            # [B] band centres of the bands held here (ascending)
    band_slice: tuple                   # (first, last+1) index of those bands in the full table
    n_bands_total: int
    power: object                       # [C, B, N]  |cwt|^2
    info: object                        # [C, B, N]  -log2(P/S + eps64)            (None if not requested)
    band_power: object                  # [C, B] fp64  sum_t P
    total_power: object                 # [C]    fp64  S (after the all-reduce when band-sharded)
    band_entropy_bits: object           # [C, B] fp64  sum_t pdf*info              (None if not requested)
    ref_bits: float                     # log2(D)/D with D = B_total * N

    def entropy_bits(self):
        """[C] total Shannon bits of the bands held here (sum over ranks when band-sharded)."""
        return self.band_entropy_bits.sum(-1)


def _host_pipelined(rt, band_order_nth, sig_wf, fs, dictionary_type, dt, host_chunks, kwargs, out_power, out_info):
    """Channel groups of a host-resident batch: H2D of group k+1 (copy stream) overlaps the kernels of group k."""
    torch = rt.torch
    x = sig_wf if isinstance(sig_wf, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(sig_wf))
    n_chan, n_points = int(x.shape[0]), int(x.shape[1])
    freq_all = scales.log_frequency_hz_from_fft_points(frequency_sample_hz=fs, fft_points=n_points,
                                                       scale_order=band_order_nth)
    bsl = kwargs["band_slice"]
    n_bands = len(freq_all) if bsl is None else int(bsl[1]) - int(bsl[0])
    if out_power is None:
        out_power = rt.empty((n_chan, n_bands, n_points), dt)
    if kwargs["want_info"] and out_info is None:
        out_info = rt.empty((n_chan, n_bands, n_points), dt)
    main = torch.cuda.current_stream(rt.device)
    copy_stream = torch.cuda.Stream(device=rt.device)
    bounds = np.linspace(0, n_chan, min(int(host_chunks), n_chan) + 1).astype(int)
    if len(bounds) > 2 and bounds[1] - bounds[0] >= 2:
        # the first group's upload is the only one that nothing overlaps: make it half a group (the rest shifts down,
        # the last group grows) -- e.g. 8 channels in 4 groups go as 1 + 2 + 2 + 3
        shift = (bounds[1] - bounds[0]) // 2
        bounds[1:-1] -= shift
    # two staging buffers owned by this call (allocated on the compute stream, so the caching allocator never has to
    # reason about the copy stream); buffer k % 2 is refilled only after the kernels of group k - 2 have consumed it
    gmax = int(np.max(np.diff(bounds)))
    stage = [torch.empty((gmax, n_points), dtype=x.dtype, device=rt.device) for _ in range(2)]
    consumed = [None, None]
    copy_stream.wait_stream(main)
    parts = []
    for k, (c0, c1) in enumerate(zip(bounds[:-1], bounds[1:])):
        buf = stage[k % 2][: c1 - c0]
        with torch.cuda.stream(copy_stream):
            if consumed[k % 2] is not None:
                copy_stream.wait_event(consumed[k % 2])
            buf.copy_(x[c0:c1], non_blocking=True)
            ready = torch.cuda.Event()
            ready.record(copy_stream)
        main.wait_event(ready)
        parts.append(cwt_power_entropy(band_order_nth, buf, fs, dictionary_type, dtype=dt, out_power=out_power[c0:c1],
                                       out_info=None if out_info is None else out_info[c0:c1], **kwargs))
        consumed[k % 2] = torch.cuda.Event()
        consumed[k % 2].record(main)
    copy_stream.wait_stream(main)
    first = parts[0]
    ent = None if first.band_entropy_bits is None else torch.cat([p.band_entropy_bits for p in parts])
    return CwtEntropy(first.frequency_hz, first.band_slice, first.n_bands_total, out_power, out_info,
                      torch.cat([p.band_power for p in parts]), torch.cat([p.total_power for p in parts]), ent,
                      first.ref_bits)


def cwt_power_entropy(band_order_nth: float, sig_wf, frequency_sample_rate_hz: float, dictionary_type: str = "norm",
                      *, dtype="float32", spectrum: str = "auto", band_slice: Optional[tuple] = None,
                      allreduce: Optional[Callable] = None, want_info: bool = True,
                      out_power=None, out_info=None, method: str = "auto",
                      truncated_bands: str = "multirate", host_chunks: int = 1) -> CwtEntropy:
    """Power, information and entropy of the order-N Gabor CWT of ``sig_wf`` ([N] or [C, N]; numpy or CUDA tensor).

    band_slice : (b0, b1) computes only bands b0..b1-1 of the standard table (band sharding of one long record).
    allreduce  : callable applied in place to the fp64 [C] tensor of local total power before normalisation
                 (e.g. ``lambda t: torch.distributed.all_reduce(t)``); None = single device.
    out_power / out_info : optional preallocated [C, B, N] device buffers to write into.
    method     : 'exact' (record FFT + per-band full-length inverse FFT, any dtype), 'multirate' (float32 only:
                 decimation pyramid + shared-memory overlap-save + half-band interpolation) or 'auto' (multirate
                 for float32 records of 2^m >= 8192 points, exact otherwise).
    truncated_bands : what the multirate method does with the lowest bands, whose atoms are cut off by the record
                 (N/s < 10).  'multirate' (default) keeps them on the fast path: the jump where the reference cuts
                 the atom is carried by exact running sums of the record inside the fused expansion (qi_mr_expand.cuh),
                 per-band power L2 against the fp64 reference 1e-5 .. 3e-5 like every other band; 'exact' recomputes
                 those bands with the full-length FFT method instead.
    host_chunks : for records that live in host memory ([C, N] numpy / CPU tensor, ideally pinned): process the
                 channels in this many groups, copying group k+1 to the device on a second stream while group k is
                 being transformed (channels are independent, so the result is identical).
    """
    rt = get_runtime(sig_wf)
    dt = dtype_name(dtype, default="float32")
    if host_chunks > 1 and getattr(rt, "name", "") == "cuda" and not getattr(sig_wf, "is_cuda", False) \
            and np.ndim(sig_wf) == 2 and np.shape(sig_wf)[0] > 1 and allreduce is None:
        return _host_pipelined(rt, band_order_nth, sig_wf, frequency_sample_rate_hz, dictionary_type, dt, host_chunks,
                               dict(spectrum=spectrum, band_slice=band_slice, want_info=want_info, method=method,
                                    truncated_bands=truncated_bands), out_power, out_info)
    sig, _ = _driver._as_2d(rt, sig_wf, dt)
    n_points = int(sig.shape[1])
    freq_all = scales.log_frequency_hz_from_fft_points(
        frequency_sample_hz=frequency_sample_rate_hz, fft_points=n_points, scale_order=band_order_nth)
    b0, b1 = (0, len(freq_all)) if band_slice is None else (int(band_slice[0]), int(band_slice[1]))
    if not (0 <= b0 < b1 <= len(freq_all)):
        raise ValueError(f"band_slice {band_slice} outside the {len(freq_all)}-band table")
    freq = freq_all[b0:b1]
    bands, scale, _, _ = _plan.gabor_bands(band_order_nth, n_points, freq, frequency_sample_rate_hz,
                                           dictionary_type, dt, spectrum)
    if method not in ("auto", "exact", "multirate"):
        raise ValueError("method must be 'auto', 'exact' or 'multirate'")
    can_mr = dt == "float32" and _plan.multirate_supported(n_points, scale)
    if method == "multirate" and not can_mr:
        raise ValueError("method='multirate' needs float32 and a record of 2^m >= 8192 points")
    deg_free = len(freq_all) * n_points
    if method != "exact" and can_mr:
        mr_bands, _, _, _ = _plan.multirate_bands(band_order_nth, n_points, freq, frequency_sample_rate_hz,
                                                  dictionary_type)
        n_trunc = int(np.count_nonzero(bands["analytic"] == 0))
        if want_info and not (truncated_bands == "exact" and n_trunc):
            # single fused pass: power plane, information plane, band sums and entropy sums written together
            res = _driver.cwt_multirate(sig, mr_bands, want_power=True, want_band_sum=True, rt=rt, out_power=out_power,
                                        want_info=True, out_info=out_info, allreduce=allreduce)
            return CwtEntropy(freq, (b0, b1), len(freq_all), res["power"], res["info"], res["band_sum"], res["total"],
                              res["entropy_sum"], float(np.log2(deg_free) / deg_free))
        res = _driver.cwt_multirate(sig, mr_bands, want_power=True, want_band_sum=True, rt=rt, out_power=out_power)
        if truncated_bands == "exact" and n_trunc:
            # the truncated atoms are the lowest-frequency rows [0, n_trunc)
            sub = _driver.cwt_fft(sig, bands[:n_trunc], frequency_sample_rate_hz, dt, want_complex=False,
                                  want_power=True, want_band_sum=True, rt=rt)
            res["power"][:, :n_trunc, :] = sub["power"]
            res["band_sum"][:, :n_trunc] = sub["band_sum"]
    else:
        res = _driver.cwt_fft(sig, bands, frequency_sample_rate_hz, dt, want_complex=False, want_power=True,
                              want_band_sum=True, rt=rt, out_power=out_power)
    power, band_power = res["power"], res["band_sum"]
    total = band_power.sum(-1)                               # [C] fp64 (tiny)
    if allreduce is not None:
        allreduce(total)
    info = ent = None
    if want_info:
        sh = _driver.shannon(power, dt, 0, total, deg_free, eps=_driver.EPS64, planes=("info",), entropy_sum=True,
                             rt=rt, out_info=out_info)
        info, ent = sh["info"], sh["entropy_sum"]
    return CwtEntropy(freq, (b0, b1), len(freq_all), power, info, band_power, total, ent,
                      float(np.log2(deg_free) / deg_free))


def stx_power_entropy(band_order_nth: float, sig_wf, frequency_sample_rate_hz: float, *, dtype="float32",
                      want_info: bool = True, out_power=None, out_info=None) -> CwtEntropy:
    """Power, information and entropy of the Stockwell transform on the standard order-N band table -- the composition

        f, t, stx = styx_stx.stx_complex_any_scale_pow2(N, sig, fs)          # styx_stx.py:195-236
        power     = np.abs(stx) ** 2
        shannon   = tfr_info.shannon_stft_from_tfr_power(power)              # tfr_info.py:231-236

    without the complex plane ever reaching HBM: the band-limited voices are interpolated straight to ``|.|^2`` with their
    fp64 band sums (csrc/qi_interp.cuh), the wide voices leave the last inverse pass the same way, and one streaming pass
    writes -log2(P / S + eps64) and the per-band entropy sums.  ``sig_wf``: [N] or [C, N], N = 2^m; float32 or float64.
    """
    rt = get_runtime(sig_wf)
    dt = dtype_name(dtype, default="float32")
    sig, _ = _driver._as_2d(rt, sig_wf, dt)
    n_points = int(sig.shape[1])
    if n_points & (n_points - 1):
        raise ValueError("the Stockwell transform needs a record of 2^m points (reference styx_stx.py:16-48 pads to one)")
    freq, bands = _plan.stx_bands(band_order_nth, n_points, frequency_sample_rate_hz)
    res = _driver.stx_fft(sig, bands, dt, want_complex=False, want_power=True, want_band_sum=True, rt=rt,
                          out_power=out_power)
    power, band_power = res["power"], res["band_sum"]
    total = band_power.sum(-1)
    deg_free = len(freq) * n_points
    info = ent = None
    if want_info:
        sh = _driver.shannon(power, dt, 0, total, deg_free, eps=_driver.EPS64, planes=("info",), entropy_sum=True, rt=rt,
                             out_info=out_info)
        info, ent = sh["info"], sh["entropy_sum"]
    return CwtEntropy(freq, (0, len(freq)), len(freq), power, info, band_power, total, ent,
                      float(np.log2(deg_free) / deg_free))
