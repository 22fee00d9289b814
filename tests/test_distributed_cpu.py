"""world_size-2 (and 3) gloo runs of the sharded hot path on the CPU tier: partitioning logic, the single
total-power all-reduce of the band-sharded case, and the collective-free channel-sharded case."""
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import qi_oracle as orc
from quantum_inferno_b200 import distributed

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def test_shard_helpers():
    assert [distributed.channel_shard(64, r, 8) for r in range(8)] == [(8 * r, 8 * r + 8) for r in range(8)]
    parts = [distributed.channel_shard(10, r, 4) for r in range(4)]
    assert parts == [(0, 3), (3, 6), (6, 8), (8, 10)]
    for nb, w in [(263, 8), (60, 8), (27, 2), (8, 8), (21, 3)]:
        edges = [distributed.band_shard(nb, r, w) for r in range(w)]
        assert edges[0][0] == 0 and edges[-1][1] == nb
        assert all(a[1] == b[0] for a, b in zip(edges, edges[1:])) and all(b1 > b0 for b0, b1 in edges)
        assert max(b1 - b0 for b0, b1 in edges) - min(b1 - b0 for b0, b1 in edges) <= 1
    # cost-weighted split balances the cost, not the count
    cost = np.r_[np.ones(10) * 9.0, np.ones(10)]
    e = [distributed.band_shard(20, r, 2, cost) for r in range(2)]
    assert e == [(0, 6), (6, 20)]
    with pytest.raises(ValueError):
        distributed.band_shard(3, 0, 4)


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_paths_gloo(tmp_path, world, golden):
    port = 29500 + world + (os.getpid() % 500)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "tests", "_dist_worker.py"), str(tmp_path)]
    env = dict(os.environ, OMP_NUM_THREADS="1", QI_EMUL_THREADS="2")
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert res.returncode == 0, res.stderr[-3000:]
    x = golden("cwt")["x2048"]
    ref = orc.cwt_power_entropy(3, x, 800.0)
    covered = []
    for rank in range(world):
        d = np.load(tmp_path / f"rank{rank}.npz")
        b0, b1 = d["band_slice"]
        covered.append((int(b0), int(b1)))
        # every rank normalises with the GLOBAL total power (the one all-reduce)
        assert abs(d["total"][0] - ref["total"]) / ref["total"] < 1e-12
        assert np.max(np.abs(d["power"][0] - ref["power"][b0:b1])) / ref["power"].max() < 1e-10
        assert np.max(np.abs(d["info"][0] - ref["info"][b0:b1])) < 1e-9
        assert abs(d["entropy_all"][0] - ref["entropy_bits"]) < 1e-10
        # channel-sharded leg
        c0, c1 = d["chan_slice"]
        batch = np.stack([x, x[::-1], np.roll(x, 7), 0.5 * x, x ** 2][:world * 2 + 1])
        for i, c in enumerate(range(c0, c1)):
            rc = orc.cwt_power_entropy(3, batch[c], 800.0)
            assert abs(d["chan_total"][i] - rc["total"]) / rc["total"] < 1e-12
            assert abs(d["chan_entropy"][i] - rc["entropy_bits"]) < 1e-10
    assert covered[0][0] == 0 and covered[-1][1] == ref["power"].shape[0]
    assert all(a[1] == b[0] for a, b in zip(covered, covered[1:]))
    # fp32 multirate, band-sharded: every rank normalised with the all-reduced total estimate
    ref3 = None
    for rank in range(world):
        d = np.load(tmp_path / f"mr_rank{rank}.npz")
        if ref3 is None:
            ref3 = orc.cwt_power_entropy(3, d["x"], 800.0)
        b0, b1 = d["band_slice"]
        assert abs(d["total"][0] - ref3["total"]) / ref3["total"] < 2e-5
        p_ref = ref3["power"][b0:b1]
        assert np.linalg.norm(d["power"][0] - p_ref) / np.linalg.norm(p_ref) < 1e-4
        strong = p_ref > 1e-2 * ref3["power"].max()
        assert np.max(np.abs(d["info"][0] - ref3["info"][b0:b1])[strong]) < 1e-3
        assert abs(d["entropy_all"][0] - ref3["entropy_bits"]) < 1e-3
