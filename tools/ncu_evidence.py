#!/usr/bin/env python3
"""Turn the raw pages of per-kernel `ncu --set full --clock-control none` captures into profiles/ (no GPU):

    python tools/ncu_evidence.py r2 gpurun_out/ev_*.raw.csv [--launches gpurun_out/ev_launches.csv]

writes profiles/<round>_kernels.md (key metrics per captured kernel, one section per case),
profiles/<round>_launches.md (launch list of the bench command) and profiles/traffic.json -- DRAM bytes
per launch keyed by case, which bench.py reports as roofline.traffic.  The captures come from
tools/prof_all.py cases run under ncu on the GPU box (tools/_evidence.sh); only the exported raw pages
travel back (a --set full report of this library is ~90 MB)."""
import collections
import csv
import json
import os
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__tex_writeback_active.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tex.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]
TRAFFIC_KEY = {"linear_named": "linear", "linear_shift": "linear_shift", "jacobian_polar": "jacobian",
               "jacobian_constj": "jacobian_constj", "wcs_batch": "wcs_batch", "apply_affine": "apply_affine"}


def to_bytes(val, unit):
    v = float(val.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def main():
    rnd = sys.argv[1]
    args = sys.argv[2:]
    launches = None
    if "--launches" in args:
        i = args.index("--launches")
        launches = args[i + 1]
        args = args[:i] + args[i + 2:]
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    traffic = {}
    with open(os.path.join(here, "profiles", rnd + "_kernels.md"), "w") as f:
        f.write("# ncu --set full --clock-control none: one launch per kernel (%s)\n\n" % rnd)
        f.write("Each section is one `tools/prof_all.py <case>` run under ncu on a B200 (8192^2 float32 frames unless "
                "the case says otherwise); the 4th launch of the kernel is captured.  Times under ncu are "
                "serialised and cold-cache: the live CUDA-event times are in the bench line.\n\n")
        for path in sorted(args):
            case = os.path.basename(path).replace("ev_", "").replace(".raw.csv", "")
            rr = list(csv.reader(open(path)))
            if len(rr) < 3:
                continue
            h, units = rr[0], rr[1]
            for r in rr[2:3]:
                name = r[h.index("Kernel Name")]
                f.write("## %s -- `%s`\n\n| metric | value | unit |\n|---|---:|---|\n" % (case, name[:110]))
                for k in KEYS:
                    if k in h:
                        f.write("| %s | %s | %s |\n" % (k, r[h.index(k)], units[h.index(k)]))
                rd = to_bytes(r[h.index("dram__bytes_read.sum")], units[h.index("dram__bytes_read.sum")])
                wr = to_bytes(r[h.index("dram__bytes_write.sum")], units[h.index("dram__bytes_write.sum")])
                f.write("\nDRAM traffic per launch: %.1f MB read + %.1f MB write = %.1f MB\n\n" % (rd / 1e6, wr / 1e6, (rd + wr) / 1e6))
                if case in TRAFFIC_KEY:
                    def pct(k):
                        try:
                            return float(r[h.index(k)].replace(",", ""))
                        except (ValueError, IndexError):
                            return None
                    traffic[TRAFFIC_KEY[case]] = {
                        "dram_bytes": rd + wr, "kernel": name[:110], "case": case,
                        "source": "profiles/%s_kernels.md" % rnd,
                        # what bounds the kernel when it is not HBM (same capture)
                        "issue_active_pct": pct("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                        "tex_writeback_pct": pct("l1tex__tex_writeback_active.avg.pct_of_peak_sustained_elapsed"),
                        "l1tex_throughput_pct": pct("l1tex__throughput.avg.pct_of_peak_sustained_elapsed"),
                        "dram_throughput_pct": pct("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                        "l1_hit_pct": pct("l1tex__t_sector_hit_rate.pct"), "l2_hit_pct": pct("lts__t_sector_hit_rate.pct")}
    with open(os.path.join(here, "profiles", "traffic.json"), "w") as f:
        json.dump(traffic, f, indent=1)
    if launches:
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
        with open(os.path.join(here, "profiles", rnd + "_launches.md"), "w") as f:
            f.write("# ncu launch list (gpu__time_duration.sum, --clock-control none)\n\n")
            f.write("Command: `python bench.py --steps 3 --warmup 3 --no-cpu --no-extra` (first 400 launches).  Per-launch "
                    "times under ncu are cold-cache and serialised: compare SHARES, not absolutes.\n\n")
            f.write("| kernel | launches | total us | mean us | share |\n|---|---:|---:|---:|---:|\n")
            for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
                f.write("| `%s` | %d | %.1f | %.1f | %.1f%% |\n" % (k[:110], len(v), sum(v), sum(v) / len(v), 100 * sum(v) / total))
            # the timed step of bench.py is two launches: the linear kernel of the named chain and the Jacobian kernel
            lin = [v for k, v in agg.items() if "k_resample_tex4<tfm::Fast, 1, 0, 0>" in k]
            jac = [v for k, v in agg.items() if "k_jac_gauss_polar_rows_pair" in k]
            if lin and jac:
                ml, mj = sum(lin[0]) / len(lin[0]), sum(jac[0]) / len(jac[0])
                f.write("\nThe timed step is one launch of each of `k_resample_tex4<Fast, 1, 0, 0>` (mean %.1f us) and "
                        "`k_jac_gauss_polar_rows_pair` (mean %.1f us): the Jacobian kernel's share of the step is %.1f %% here; "
                        "the bench line's live `roofline.share_of_step` must agree (the other launches of the list are the "
                        "parity-mode step, the end-to-end legs and the A/B variants of the breakdown).\n"
                        % (ml, mj, 100 * mj / (ml + mj)))
    print(json.dumps(traffic, indent=1))


if __name__ == "__main__":
    main()
