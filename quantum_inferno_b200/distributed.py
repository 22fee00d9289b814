"""
Multi-GPU partitioning of the hot path: one process per GPU, ``torch.distributed`` (NCCL over NVLink) for the
plumbing.

* Records (channels) are independent in every transform and every ``tfr_info`` reduction is per record
  (reference tfr_info.py:236,247,259), so a batch shards by channel with NO data-path collective.
* One long record can instead be sharded by band (bands never interact: styx_cwt.py:195 is row-wise,
  styx_stx.py:231-234 loops bands).  The only exchange is then ONE all-reduce (sum, fp64, one scalar per
  record) of the total power S that normalises the global pdf of ``shannon_stft_from_tfr_power``.
"""
from typing import Optional, Sequence, Tuple

import numpy as np


def channel_shard(n_channels: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [c0, c1) of records for this rank (sizes differ by at most one)."""
    base, extra = divmod(int(n_channels), int(world))
    c0 = rank * base + min(rank, extra)
    return c0, c0 + base + (1 if rank < extra else 0)


def band_shard(n_bands: int, rank: int, world: int, cost: Optional[Sequence[float]] = None) -> Tuple[int, int]:
    """Contiguous band range [b0, b1) for this rank, balanced on ``cost`` (default: equal cost per band, which is
    what the full-length inverse FFT per band costs).  Every rank gets at least one band when n_bands >= world."""
    n_bands, world = int(n_bands), int(world)
    if n_bands < world:
        raise ValueError(f"cannot shard {n_bands} bands over {world} ranks")
    w = np.ones(n_bands) if cost is None else np.asarray(cost, dtype=np.float64)
    cum = np.concatenate(([0.0], np.cumsum(w)))
    edges = [0]
    for r in range(1, world):
        target = cum[-1] * r / world
        e = int(np.searchsorted(cum, target, side="left"))
        e = max(e, edges[-1] + 1)                       # at least one band per rank
        e = min(e, n_bands - (world - r))               # leave one for every later rank
        edges.append(e)
    edges.append(n_bands)
    return edges[rank], edges[rank + 1]


def sum_allreduce(group=None):
    """Returns the callable ``cwt_entropy.cwt_power_entropy(allreduce=...)`` expects: an in-place SUM all-reduce
    of the fp64 per-record total power over ``group`` (NCCL for CUDA tensors, gloo for host buffers)."""
    import torch
    import torch.distributed as dist

    def _reduce(total):
        t = total if isinstance(total, torch.Tensor) else torch.from_numpy(total)    # shares memory with numpy
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        _reduce.calls += 1
        return total

    _reduce.calls = 0            # collectives issued through this callable (bench.py reports it)
    return _reduce


def cwt_power_entropy_band_sharded(band_order_nth, sig_wf, frequency_sample_rate_hz, rank=None, world=None,
                                   group=None, allreduce=None, **kwargs):
    """Band-sharded ``cwt_entropy.cwt_power_entropy``: every rank holds the whole record(s), computes its own band
    range and joins the single total-power all-reduce (``allreduce``: the callable to use, default
    ``sum_allreduce(group)``).  Returns this rank's ``CwtEntropy`` (its bands only)."""
    import torch.distributed as dist
    from . import cwt_entropy, scales_dyadic as scales
    rank = dist.get_rank(group) if rank is None else rank
    world = dist.get_world_size(group) if world is None else world
    n_points = int(sig_wf.shape[-1])
    freq = scales.log_frequency_hz_from_fft_points(frequency_sample_rate_hz, n_points, band_order_nth)
    return cwt_entropy.cwt_power_entropy(band_order_nth, sig_wf, frequency_sample_rate_hz,
                                         band_slice=band_shard(len(freq), rank, world,
                                                               band_cost(band_order_nth, n_points, freq,
                                                                         frequency_sample_rate_hz, kwargs)),
                                         allreduce=allreduce if allreduce is not None else sum_allreduce(group),
                                         **kwargs)


# measured cost of one band of the fused fp32 multirate path relative to a deep band (B200, 2^24 samples): the full
# rate convolution + separate information pass of a level-0 band, the larger input share of the x2 / x4 interpolators
_MULTIRATE_LEVEL_COST = {0: 3.6, 1: 3.8, 2: 2.5, 3: 1.5, 4: 1.25}


def band_cost(band_order_nth, n_points, frequency_hz, frequency_sample_rate_hz, kwargs=None):
    """Per-band cost estimate for ``band_shard``: equal for the exact method (one full-length inverse FFT per band),
    level dependent for the multirate method."""
    from . import _plan
    kwargs = kwargs or {}
    dt = str(kwargs.get("dtype", "float32")).replace("torch.", "")
    if dt != "float32" or kwargs.get("method", "auto") == "exact":
        return None
    bands, scale, _, _ = _plan.multirate_bands(band_order_nth, n_points, frequency_hz, frequency_sample_rate_hz,
                                               kwargs.get("dictionary_type", "norm"))
    if not _plan.multirate_supported(n_points, scale):
        return None
    return [_MULTIRATE_LEVEL_COST.get(int(lv), 1.05) for lv in bands["level"]]
