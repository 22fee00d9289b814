"""Worker for tests/test_distributed_cpu.py: one gloo rank driving the band-/channel-sharded hot path with the
test-only emulator runtime (no GPU in the CPU test tier)."""
import os
import sys

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from quantum_inferno_b200 import _runtime, cwt_entropy, distributed
    from tests.emul.emul_runtime import EmulRuntime

    out_dir = sys.argv[1]
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    fs, n = 800.0, 2048
    g = np.load(os.path.join(ROOT, "tests", "golden", "cwt.npz"))
    x = g["x2048"]
    with _runtime.use_runtime(EmulRuntime()):
        # (1) one long record, band-sharded, single total-power all-reduce
        r = distributed.cwt_power_entropy_band_sharded(3, x, fs, dtype="float64")
        # (2) channel-sharded batch: no collective on the data path
        batch = np.stack([x, x[::-1], np.roll(x, 7), 0.5 * x, x ** 2][:world * 2 + 1])
        c0, c1 = distributed.channel_shard(len(batch), rank, world)
        rc = cwt_entropy.cwt_power_entropy(3, batch[c0:c1], fs, dtype="float64")
        # (3) the fp32 multirate path, band-sharded: estimate -> ONE all-reduce of the total -> fused expansion
        n3 = 8192
        k = np.arange(n3)
        x3 = np.cos(2 * np.pi * 60.0 / fs * k) + 0.3 * np.cos(2 * np.pi * 7.0 / fs * k) \
            + np.random.default_rng(11).standard_normal(n3) / 16.0
        r3 = distributed.cwt_power_entropy_band_sharded(3, x3, fs, dtype="float32", method="multirate")
    ent3 = torch.tensor([float(np.sum(r3.band_entropy_bits))], dtype=torch.float64)
    dist.all_reduce(ent3)
    np.savez(os.path.join(out_dir, f"mr_rank{rank}.npz"), band_slice=np.array(r3.band_slice), power=r3.power,
             info=r3.info, total=r3.total_power, entropy_all=ent3.numpy(), x=x3)
    ent = torch.tensor([float(np.sum(r.band_entropy_bits))], dtype=torch.float64)
    dist.all_reduce(ent)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), band_slice=np.array(r.band_slice), power=r.power, info=r.info,
             total=r.total_power, band_entropy=r.band_entropy_bits, entropy_all=ent.numpy(),
             chan_slice=np.array([c0, c1]), chan_total=rc.total_power, chan_entropy=rc.band_entropy_bits.sum(-1))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
