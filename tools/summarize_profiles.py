"""
Turns the ncu captures a gpurun call brought back (gpurun_out/) into the small text artefacts committed under
profiles/:

    python tools/summarize_profiles.py launches gpurun_out/r01_multirate_launches.csv profiles/r01_multirate_launch_summary.md
    python tools/summarize_profiles.py full     gpurun_out/r01_multirate_full.ncu-rep profiles/r01_multirate_full_summary.md [profiles/traffic.json]

`launches` reads the CSV of `ncu --metrics gpu__time_duration.sum --csv`; `full` reads a `--set full` report through
`ncu -i ... --page raw --csv` (ncu must be on PATH; no GPU needed).
"""
import collections
import csv
import io
import json
import re
import subprocess
import sys


def short(name):
    name = re.sub(r"\(.*", "", name)
    name = name.replace("void ", "").replace("qi::", "")
    return name


def launches(src, dst):
    rows = [r for r in csv.reader(open(src)) if len(r) > 10]
    hdr, rows = rows[0], rows[1:]
    ki, vi, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
    mine = [(short(r[ki]), r[gi], float(r[vi]) / 1e6) for r in rows if "qi::" in r[ki]]
    starts = [i for i, m in enumerate(mine) if m[0].startswith("mr_table_kernel") or m[0].startswith("fft_pass")]
    per = collections.OrderedDict()
    for n, g, ms in mine:
        per.setdefault(n, [0, 0.0])
        per[n][0] += 1
        per[n][1] += ms
    tot = sum(v[1] for v in per.values())
    with open(dst, "w") as f:
        f.write(f"# ncu launch list summary ({src.split('/')[-1]})\n\n")
        f.write("`ncu --metrics gpu__time_duration.sum --clock-control none` over `python bench.py --steps 2 --warmup 3`; "
                "per-launch times are serialised and cold-cache, so the SHARES are what compares with the live "
                "CUDA-event categories in the bench JSON.\n\n")
        f.write("| kernel | launches | total ms | share |\n|---|---|---|---|\n")
        for n, (cnt, ms) in sorted(per.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{n}` | {cnt} | {ms:.3f} | {100 * ms / tot:.1f} % |\n")
        if len(starts) >= 2:
            # one resident step = the run of launches between two table kernels with the largest total time
            # (the end-to-end leg of the bench re-runs the step per channel group with smaller grids)
            spans = list(zip(starts[:-1], starts[1:]))
            s, e = max(spans, key=lambda se: sum(m[2] for m in mine[se[0]:se[1]]))
            f.write("\n## one step (launch order)\n\n| kernel | grid | ms |\n|---|---|---|\n")
            step_tot = 0.0
            for n, g, ms in mine[s:e]:
                step_tot += ms
                if ms >= 0.02:
                    f.write(f"| `{n}` | {g} | {ms:.3f} |\n")
            f.write(f"\nsum of the step's launches: {step_tot:.2f} ms\n")
    print("wrote", dst)


FULL_METRICS = [
    ("gpu__time_duration.sum", "time"),
    ("smsp__inst_executed.sum", "warp instr"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %"),
    ("launch__registers_per_thread", "regs"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem pipe %"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma pipe %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu pipe %"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_sb"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_sb"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall mio"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait"),
]


def full(src, dst, traffic_json=None):
    if src.endswith(".csv"):      # already exported on the GPU box (`ncu -i report --page raw --csv`): reports are too big to bring back
        out = open(src).read()
    else:
        out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    have = [(m, lab) for m, lab in FULL_METRICS if m in col]
    traffic = {}
    with open(dst, "w") as f:
        f.write(f"# ncu --set full summary ({src.split('/')[-1]})\n\n")
        f.write("`ncu --set full --clock-control none --import-source on` on one step of the headline workload "
                "(`tools/profile_step.py`); one row per captured launch.\n\n")
        f.write("| kernel | grid | " + " | ".join(f"{lab} [{units[col[m]]}]" if units[col[m]] else lab for m, lab in have) + " |\n")
        f.write("|---|---|" + "---|" * len(have) + "\n")
        for r in data:
            name = short(r[col["Kernel Name"]])
            vals = []
            for m, _ in have:
                v = r[col[m]]
                try:
                    v = f"{float(v):.4g}"
                except ValueError:
                    pass
                vals.append(v)
            f.write(f"| `{name}` | {r[col['Grid Size']]} | " + " | ".join(vals) + " |\n")
            if "dram__bytes_read.sum" in col:
                def to_bytes(metric):
                    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[units[col[metric]]]
                    return float(r[col[metric]]) * scale
                key = f"{name} {r[col['Grid Size']]}"
                traffic[key] = {"dram_bytes": to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum"),
                                "ms": float(r[col["gpu__time_duration.sum"]])}
    print("wrote", dst)
    if traffic_json:
        json.dump(traffic, open(traffic_json, "w"), indent=1)
        print("wrote", traffic_json)


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        full(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else None)
