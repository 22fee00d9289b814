"""
bench.py -- headline benchmark of the quantum-inferno B200 time-frequency path.

Workload (BASELINE.json north_star): fused order-3 Gabor CWT + power + Shannon information/entropy, fp32,
CH_PER_GPU channels x 2^24 samples @ 800 Hz per GPU (60 bands) -- 64 x 2^24 on 8 GPUs, weak scaling by
channel with no data-path collective.  A step is one pass of the hot path over that resident batch.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference ...                            # CPU port of the reference (oracle), host cores
    torchrun --nproc-per-node N bench.py --gpus N ...               # one rank per GPU

Rank 0 prints ONE JSON line (contract in the task statement): value = TFR cells/s over all GPUs with inputs
resident in HBM; e2e = same metric through the public API from pinned host memory incl. H2D of the records and
D2H of the entropy summaries; roofline = dominant kernel vs the measured HBM copy peak; cpu_baseline = the
oracle timed on the host.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FS = 800.0
ORDER = 3
LOG2_N = int(os.environ.get("QI_BENCH_LOG2N", "24"))
CH_PER_GPU = int(os.environ.get("QI_BENCH_CHANNELS", "8"))
ALG_BYTES_PER_CELL_F32 = 2 * 4 + 4.0 / 60.0       # power + info planes + input share
HOST_CHUNKS = int(os.environ.get("QI_BENCH_HOST_CHUNKS", "4"))
METHOD = os.environ.get("QI_BENCH_METHOD", "multirate")     # 'multirate' (default fast path) or 'exact'
ALGORITHMS = {
    "multirate": "multirate fp32 path: half-band pyramid, per-level overlap-save FFT convolution in shared memory "
                 "(level 0 -> power rows), polyphase interpolation fused with |.|^2, -log2(P/S+eps), both plane stores "
                 "and the band/entropy sums; one streaming information pass for the level-0 rows",
    "exact": "exact path: record FFT + per-band 3-pass inverse FFT through HBM"}
KERNEL_OF = {
    "multirate": {"fft_fwd": "mr_table_kernel+mr_decimate_kernel",
                  "inv_first": "mr_level2k_kernel[levels>=1]+mr_expand_kernel<MID>",
                  "inv_mid": "mr_level2k_kernel[level 0]", "inv_last": "mr_expand_kernel<POWER_INFO>",
                  "info": "mr_info_rows_kernel"},
    "exact": {"fft_fwd": "fft_pass_kernel[forward]", "inv_first": "fft_pass_kernel[spectrum x response]",
              "inv_mid": "fft_pass_kernel[middle]", "inv_last": "fft_pass_kernel[slice + power]",
              "info": "shannon_kernel"}}
METRIC = "tfr_cells_per_s"
UNIT = "cells/s"


def workload_name():
    return (f"north_star: N=3 Gabor CWT + power + info + entropy, {CH_PER_GPU} ch/GPU x 2^{LOG2_N} samples @ 800 Hz, "
            f"fp32 planes")


# ----------------------------------------------------------------------------- synthetic input
def synth_channel_numpy(n, chan):
    k = np.arange(n, dtype=np.float64)
    f_c = 60.0 * 2.0 ** ((chan % 12) / 12.0)
    dur = n / FS
    chirp = 0.5 * np.cos(2 * np.pi * (1.0 * k / FS + 0.5 * (199.0 / dur) * (k / FS) ** 2))
    noise = np.random.default_rng(1234 + chan).standard_normal(n) * 2.0 ** -4
    return np.cos(2 * np.pi * f_c / FS * k) + chirp + noise


def synth_batch_torch(torch, n, chans, device):
    out = torch.empty(len(chans), n, dtype=torch.float32, device=device)
    k = torch.arange(n, dtype=torch.float64, device=device)
    dur = n / FS
    for i, c in enumerate(chans):
        f_c = 60.0 * 2.0 ** ((c % 12) / 12.0)
        g = torch.Generator(device=device).manual_seed(1234 + c)
        x = torch.cos(2 * np.pi * f_c / FS * k) + 0.5 * torch.cos(2 * np.pi * (k / FS + 0.5 * (199.0 / dur) * (k / FS) ** 2))
        x += torch.randn(n, dtype=torch.float64, device=device, generator=g) * 2.0 ** -4
        out[i] = x.float()
    return out


# ----------------------------------------------------------------------------- CPU port (oracle) timing
def _oracle_one(args):
    chan, n = args
    from oracle import qi_oracle as orc
    x = synth_channel_numpy(n, chan)
    r = orc.cwt_power_entropy(ORDER, x, FS)
    return r["power"].shape[0] * n, float(r["entropy_bits"])


def cpu_baseline_single(log2n=None):
    """~10-30 s of single-threaded CPU work on a bounded sample of the same workload."""
    log2n = int(os.environ.get("QI_BENCH_CPU_LOG2N", "20")) if log2n is None else log2n
    n = 1 << log2n
    t0 = time.perf_counter()
    cells, _ = _oracle_one((0, n))
    dt = time.perf_counter() - t0
    return {"value": cells / dt, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"oracle/qi_oracle.cwt_power_entropy (numpy fp64 restatement of styx_cwt + tfr_info), channel 0, "
                      f"first 2^{log2n} samples ({cells // n} bands), {dt:.1f} s",
            "samples_per_s": n / dt}


def run_reference_arm(args):
    """bench.py --impl reference: the reference's CPU algorithm (numpy port; the Python reference itself cannot
    travel to the GPU box) on all host cores, one process per channel, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = max(1, min(os.cpu_count() or 1, int(os.environ.get("QI_BENCH_CPU_PROCS", "64"))))
    log2n = int(os.environ.get("QI_BENCH_REF_LOG2N", "18"))
    n = 1 << log2n
    jobs = [(c, n) for c in range(cores)]
    with mp.get_context("fork").Pool(cores) as pool:
        for _ in range(args.warmup):
            pool.map(_oracle_one, jobs[:cores])
        t0 = time.perf_counter()
        cells = 0
        for _ in range(args.steps):
            cells += sum(c for c, _ in pool.map(_oracle_one, jobs))
        dt = time.perf_counter() - t0
    value = cells / dt
    sample = (f"{cores} processes x 1 channel x 2^{log2n} samples per step (numpy fp64 port of the reference, "
              f"{cells // (args.steps * cores * n)} bands)")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(), "sample": sample},
            "samples_per_s": cores * n * args.steps / dt,
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.path = tempfile.mktemp(prefix="qi_clocks_", suffix=".csv")
        self.proc = None
        self.idx = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "25"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in open(self.path):
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1])); mx.append(float(parts[2])); power.append(float(parts[3]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        os.unlink(self.path)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(power)),
                "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------- our arm
def run_gpu_arm(args):
    # the contract is ONE JSON line on stdout: route anything libraries print to fd 1 (e.g. the NCCL banner) to stderr
    json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    import torch
    from quantum_inferno_b200 import _lib, _runtime, cwt_entropy

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device; there is no CPU fallback (use --impl reference for the CPU port)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=dev)
    rt = _runtime.get_runtime()
    lib = rt.lib

    n = 1 << LOG2_N
    chans = [rank * CH_PER_GPU + i for i in range(CH_PER_GPU)]
    x = synth_batch_torch(torch, n, chans, dev)                       # resident input, 4*C*N bytes (> L2)
    freq = cwt_entropy.scales.log_frequency_hz_from_fft_points(FS, n, ORDER)
    n_bands = len(freq)
    from quantum_inferno_b200 import _plan
    n_level0 = int(np.count_nonzero(_plan.multirate_bands(ORDER, n, freq, FS, "norm")[0]["level"] == 0))
    cells_per_step_gpu = CH_PER_GPU * n_bands * n
    power = torch.empty(CH_PER_GPU, n_bands, n, dtype=torch.float32, device=dev)
    info = torch.empty_like(power)

    def step(src):
        return cwt_entropy.cwt_power_entropy(ORDER, src, FS, dtype="float32", out_power=power, out_info=info,
                                             method=METHOD)

    def sync_all():
        torch.cuda.synchronize(dev)
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize(dev)

    for _ in range(args.warmup):
        r = step(x)
    sync_all()

    # ---- timed region: K steps, inputs resident in HBM
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    lib.qi_profile_enable(1)
    launches0 = lib.qi_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    e0.record()
    for _ in range(args.steps):
        r = step(x)
    e1.record()
    sync_all()
    elapsed_ms = e0.elapsed_time(e1)
    launches = lib.qi_launch_count() - launches0
    lib.qi_profile_enable(0)
    cat_ms = (np.zeros(_lib.QI_N_CATEGORIES), np.zeros(_lib.QI_N_CATEGORIES, dtype=np.int64))
    lib.qi_profile_read(cat_ms[0].ctypes.data, cat_ms[1].ctypes.data)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms = float(t.item())
    entropy_check = float(r.entropy_bits()[0].item())

    # ---- end to end through the public API: pinned host records -> H2D -> kernels -> D2H of the summaries
    x_host = torch.empty(CH_PER_GPU, n, dtype=torch.float32, pin_memory=True)
    x_host.copy_(x)
    e2e_steps = max(1, min(args.steps, 5))
    def step_host(src):
        # the call a user makes for host-resident records: channel groups, H2D of group k+1 under the kernels of group k
        return cwt_entropy.cwt_power_entropy(ORDER, src, FS, dtype="float32", out_power=power, out_info=info,
                                             method=METHOD, host_chunks=HOST_CHUNKS)

    step_host(x_host)
    sync_all()
    # the box's plain pinned-host -> HBM bandwidth for the same buffer: the floor of any end-to-end number
    xd_probe = torch.empty_like(x)
    e0.record()
    xd_probe.copy_(x_host, non_blocking=True)
    e1.record()
    sync_all()
    h2d_gbps = x_host.numel() * 4 / (e0.elapsed_time(e1) * 1e-3) / 1e9
    del xd_probe
    e0.record()
    d2h = 0
    for _ in range(e2e_steps):
        rr = step_host(x_host)
        outs = [rr.band_entropy_bits.cpu(), rr.band_power.cpu(), rr.total_power.cpu()]
        d2h = sum(o.numel() * o.element_size() for o in outs)
    e1.record()
    sync_all()
    t2 = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
    e2e_ms = float(t2.item()) / e2e_steps

    if rank == 0:
        ms_per_step = elapsed_ms / args.steps
        value = world * cells_per_step_gpu / (ms_per_step * 1e-3)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except (OSError, ValueError):
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "6650 GB/s (of fallback)"
        # dominant kernel = the category with the largest summed device time in the timed region
        dom = int(np.argmax(cat_ms[0]))
        dom_launches = int(cat_ms[1][dom])
        dom_avg_ms = float(cat_ms[0][dom] / max(1, dom_launches))
        # cells the dominant category's launches process per step: the expand launches of the multirate path write the
        # bands of level >= 1 only (the level-0 rows are written by the level-0 convolution + the information pass)
        dom_cells_step = cells_per_step_gpu
        if METHOD == "multirate" and _lib.CATEGORY_NAMES[dom] == "inv_last":
            dom_cells_step = CH_PER_GPU * (n_bands - n_level0) * n
        cells_per_launch = dom_cells_step * args.steps / max(1, dom_launches)
        achieved = ALG_BYTES_PER_CELL_F32 * cells_per_launch / (dom_avg_ms * 1e-3) / 1e9
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))[METHOD].get(_lib.CATEGORY_NAMES[dom])
        except (OSError, ValueError):
            pass
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(), "channels_total": world * CH_PER_GPU, "bands": n_bands,
                       "parallelism": f"channel-sharded x{world}, no data-path collective",
                       "l2": "inputs (0.5 GB) and planes (64 GB) per step are far larger than L2; no flush needed",
                       "algorithm": ALGORITHMS[METHOD]},
            "samples_per_s": world * CH_PER_GPU * n / (ms_per_step * 1e-3),
            "roofline": {"bound": "hbm", "kernel": KERNEL_OF[METHOD].get(_lib.CATEGORY_NAMES[dom], _lib.CATEGORY_NAMES[dom]),
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src,
                         "avg_launch_ms": dom_avg_ms, "launches": dom_launches,
                         "cells_per_launch": cells_per_launch,
                         "alg_bytes_per_cell": ALG_BYTES_PER_CELL_F32,
                         "step_frac": ALG_BYTES_PER_CELL_F32 * cells_per_step_gpu / (ms_per_step * 1e-3) / 1e9 / peak,
                         "category_ms_per_step": {nm: float(ms) / args.steps for nm, ms in zip(_lib.CATEGORY_NAMES, cat_ms[0])}},
            "cpu_baseline": cpu_baseline_single() if world == 1 else None,
            "e2e": {"value": world * cells_per_step_gpu / (e2e_ms * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": int(x_host.numel() * 4), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": e2e_ms, "h2d_gbps_measured": h2d_gbps,
                    "h2d_floor_ms": x_host.numel() * 4 / (h2d_gbps * 1e9) * 1e3,
                    "note": f"public API cwt_entropy.cwt_power_entropy(host_chunks={HOST_CHUNKS}) on pinned host records (H2D of the "
                            "next channel group overlaps the kernels of the current one); planes stay in HBM, "
                            "entropy/power summaries are read back"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "check": {"entropy_bits_ch0": entropy_check},
        }
        json_out.write(json.dumps(line) + "\n")
        json_out.flush()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
