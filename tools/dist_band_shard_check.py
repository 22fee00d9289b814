"""
Band-sharded fused CWT + entropy over NCCL on real GPUs (SURVEY 8e, config 5 shape scaled down): every rank holds
the record, computes its band range and joins ONE fp64 all-reduce of the total power.  Checks the result against the
single-GPU call on rank 0 and reports device times (max over ranks).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/dist_band_shard_check.py [log2n] [order]
"""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from bench import FS, synth_batch_torch  # noqa: E402
from quantum_inferno_b200 import cwt_entropy, distributed  # noqa: E402


def main():
    log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    order = float(sys.argv[2]) if len(sys.argv) > 2 else 12.0
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    rank, world = dist.get_rank(), dist.get_world_size()
    n = 1 << log2n
    x = synth_batch_torch(torch, n, [0], dev)                      # the same record on every rank

    # this rank's planes are allocated once and reused (at config-5 size they are most of the GPU)
    from quantum_inferno_b200 import scales_dyadic as sc
    freq = sc.log_frequency_hz_from_fft_points(FS, n, order)
    b0, b1 = distributed.band_shard(len(freq), rank, world, distributed.band_cost(order, n, freq, FS, {"dtype": "float32"}))
    power = torch.empty(1, b1 - b0, n, dtype=torch.float32, device=dev)
    info = torch.empty_like(power)

    def sharded():
        return distributed.cwt_power_entropy_band_sharded(order, x, FS, dtype="float32", out_power=power, out_info=info)

    r = sharded()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 3
    e0.record()
    for _ in range(reps):
        r = sharded()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ent = r.band_entropy_bits.sum(-1).clone()                      # this rank's bands
    dist.all_reduce(ent)
    pdf = (r.band_power.sum(-1) / r.total_power).reshape(1)       # this rank's share of the pdf (band sums are fp64)
    dist.all_reduce(pdf)
    b0, b1 = r.band_slice
    out = {"world": world, "log2n": log2n, "order": order, "bands_total": r.n_bands_total,
           "ms_per_call_max_over_ranks": float(t.item()), "entropy_bits_allranks": float(ent[0].item()),
           "pdf_sum_allranks": float(pdf.item())}
    fits = r.n_bands_total * n * 8 < 60e9                       # both planes of ALL bands on one GPU next to this rank's
    if rank == 0 and not fits:
        out["cells_per_s"] = r.n_bands_total * n / (out["ms_per_call_max_over_ranks"] * 1e-3)
        out["note"] = "single-GPU comparison skipped: the full planes do not fit next to this rank's share"
        print(json.dumps(out), flush=True)
        assert abs(out["pdf_sum_allranks"] - 1.0) < 1e-5
    if rank == 0 and fits:
        full = cwt_entropy.cwt_power_entropy(order, x, FS, dtype="float32")
        out["entropy_bits_single_gpu"] = float(full.entropy_bits()[0].item())
        out["total_power_rel_diff"] = abs(float(full.total_power[0]) - float(r.total_power[0])) / float(full.total_power[0])
        out["rank0_power_max_abs_diff"] = float((full.power[:, b0:b1] - r.power).abs().max().item())
        out["rank0_info_max_abs_diff"] = float((full.info[:, b0:b1] - r.info).abs().max().item())
        cells = r.n_bands_total * n
        out["cells_per_s"] = cells / (out["ms_per_call_max_over_ranks"] * 1e-3)
        print(json.dumps(out), flush=True)
        assert abs(out["entropy_bits_allranks"] - out["entropy_bits_single_gpu"]) < 1e-6
        assert abs(out["pdf_sum_allranks"] - 1.0) < 1e-5 and out["total_power_rel_diff"] < 1e-8
        assert out["rank0_power_max_abs_diff"] == 0.0 and out["rank0_info_max_abs_diff"] < 1e-5
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
