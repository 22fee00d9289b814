"""Timeline of the host-pipelined public-API call (pinned host records -> H2D -> kernels): event timestamps of every
group's copy and transform, on the default stream and on a side stream.   python tools/e2e_probe.py [chunks]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from bench import FS, ORDER, synth_batch_torch  # noqa: E402
from quantum_inferno_b200 import cwt_entropy  # noqa: E402

dev = torch.device("cuda", 0)
n = 1 << 24
chunks = int(sys.argv[1]) if len(sys.argv) > 1 else 4
x = synth_batch_torch(torch, n, list(range(8)), dev)
nb = len(cwt_entropy.scales.log_frequency_hz_from_fft_points(FS, n, ORDER))
power = torch.empty(8, nb, n, dtype=torch.float32, device=dev)
info = torch.empty_like(power)
xh = torch.empty(8, n, dtype=torch.float32, pin_memory=True)
xh.copy_(x)


def timeline(main):
    """Hand-rolled copy of cwt_entropy._host_pipelined with timing events."""
    cs = torch.cuda.Stream()
    bounds = np.linspace(0, 8, chunks + 1).astype(int)
    stage = [torch.empty((int(np.max(np.diff(bounds))), n), dtype=torch.float32, device=dev) for _ in range(2)]
    consumed = [None, None]
    ev = []
    torch.cuda.synchronize()
    t_host0 = time.perf_counter()
    origin = torch.cuda.Event(enable_timing=True)
    origin.record(main)
    cs.wait_stream(main)
    host_marks = []
    for k, (c0, c1) in enumerate(zip(bounds[:-1], bounds[1:])):
        buf = stage[k % 2][: c1 - c0]
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        with torch.cuda.stream(cs):
            if consumed[k % 2] is not None:
                cs.wait_event(consumed[k % 2])
            e[0].record(cs)
            buf.copy_(xh[c0:c1], non_blocking=True)
            e[1].record(cs)
        main.wait_event(e[1])
        e[2].record(main)
        cwt_entropy.cwt_power_entropy(ORDER, buf, FS, dtype="float32", out_power=power[c0:c1], out_info=info[c0:c1])
        e[3].record(main)
        consumed[k % 2] = e[3]
        ev.append(e)
        host_marks.append(time.perf_counter() - t_host0)
    torch.cuda.synchronize()
    for k, e in enumerate(ev):
        print(f"  group {k}: copy [{origin.elapsed_time(e[0]):6.2f}, {origin.elapsed_time(e[1]):6.2f}]  transform "
              f"[{origin.elapsed_time(e[2]):6.2f}, {origin.elapsed_time(e[3]):6.2f}] ms   host enqueue done at "
              f"{host_marks[k] * 1e3:6.2f} ms", flush=True)


print("default stream:")
timeline(torch.cuda.current_stream())
timeline(torch.cuda.current_stream())
side = torch.cuda.Stream()
print("side stream:")
with torch.cuda.stream(side):
    timeline(side)
    timeline(side)
