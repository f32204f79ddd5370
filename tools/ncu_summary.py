#!/usr/bin/env python3
"""Summarise ncu outputs into profiles/ (run in the authoring container, no GPU needed).

    python tools/ncu_summary.py gpurun_out/launches_r1.csv gpurun_out/prof_r1.ncu-rep profiles/r1

writes <prefix>_launches.md (per-kernel launch count / total / mean device time and share of the
step), <prefix>_kernels.md (key metrics of the --set full capture) and profiles/traffic.json (DRAM
bytes per launch of the two hot kernels, consumed by bench.py's roofline.traffic)."""
import collections
import csv
import json
import os
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]


def to_bytes(val, unit):
    v = float(val.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def main():
    launches, rep, prefix = sys.argv[1:4]
    rows = [r for r in csv.reader(open(launches)) if len(r) > 5]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1e-3)
        agg.setdefault(r[ki], []).append(v)
    total = sum(sum(v) for v in agg.values())
    with open(prefix + "_launches.md", "w") as f:
        f.write("# ncu launch list (gpu__time_duration.sum, --clock-control none)\n\n")
        f.write("Command: `python bench.py --steps 3 --warmup 3 --no-cpu` (first 160 launches).  Per-launch "
                "times under ncu are cold-cache and serialised: compare SHARES, not absolutes.\n\n")
        f.write("| kernel | launches | total us | mean us | share |\n|---|---:|---:|---:|---:|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            f.write("| `%s` | %d | %.1f | %.1f | %.1f%% |\n" % (k[:110], len(v), sum(v), sum(v) / len(v), 100 * sum(v) / total))
    if rep.endswith(".csv"):      # `ncu -i X.ncu-rep --page raw --csv` already run on the GPU box (big reports do not travel)
        raw = open(rep).read()
    else:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    h, units = rr[0], rr[1]
    traffic = {}
    with open(prefix + "_kernels.md", "w") as f:
        f.write("# ncu --set full --clock-control none (one launch per kernel, timed-region launches)\n\n")
        for r in rr[2:]:
            name = r[h.index("Kernel Name")]
            f.write("## `%s`\n\n| metric | value | unit |\n|---|---:|---|\n" % name[:120])
            for k in KEYS:
                if k in h:
                    f.write("| %s | %s | %s |\n" % (k, r[h.index(k)], units[h.index(k)]))
            rd = to_bytes(r[h.index("dram__bytes_read.sum")], units[h.index("dram__bytes_read.sum")])
            wr = to_bytes(r[h.index("dram__bytes_write.sum")], units[h.index("dram__bytes_write.sum")])
            f.write("\nDRAM traffic per launch: %.1f MB read + %.1f MB write = %.1f MB\n\n" % (rd / 1e6, wr / 1e6, (rd + wr) / 1e6))
            key = None
            if "k_jacobian" in name:
                key = "jacobian"
            elif "k_resample_tex" in name or "k_resample<(int)1" in name or "k_resample<1" in name:
                key = "linear"
            if key:
                traffic[key] = rd + wr
    with open(os.path.join(os.path.dirname(prefix), "traffic.json"), "w") as f:
        json.dump(traffic, f, indent=1)
    print(open(prefix + "_launches.md").read())
    print(traffic)


if __name__ == "__main__":
    main()
