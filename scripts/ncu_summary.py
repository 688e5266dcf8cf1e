#!/usr/bin/env python
"""Summarise an ncu raw-page CSV (ncu -i X.ncu-rep --page raw --csv) into the handful of
numbers the roofline discussion needs.  Usage: ncu_summary.py raw.csv [out.md]"""
import csv
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 % of peak"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_active", "L1/tex % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM % of peak"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64 pipe %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu pipe %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu pipe %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu pipe %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma pipe %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "stall long_scoreboard"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_sb /issue"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall lg_throttle /issue"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_throttle /issue"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall mio_throttle /issue"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_sb /issue"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier /issue"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait /issue"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall not_selected /issue"),
    ("smsp__average_warps_issue_stalled_tex_throttle_per_issue_active.ratio", "stall tex_throttle /issue"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "L1 global load sectors"),
    ("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "L1 global load requests"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("smsp__inst_executed.sum", "warp instructions"),
]


SHORT = [("bp_correct_kernel", "bp_correct"), ("gauss_tile_kernel", "gaussian_u16_f32"), ("translate_u16_rows_kernel", "translate_u16"), ("translate_u16_tma_kernel", "translate_u16"),
         ("split_stats_kernel", "precode_delta_split_stats"), ("delta_split_kernel", "precode_delta_split"),
         ("movie_stats_kernel", "stats_minmax_hist")]


def to_bytes(value, unit):
    v = float(value.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[unit]


def launch_shares(path):
    """`ncu --metrics gpu__time_duration.sum --csv` launch list -> per-kernel totals and shares."""
    lines = [ln for ln in open(path) if ln.startswith('"')]
    rows = list(csv.reader(lines))
    hdr, data = rows[0], rows[1:]
    idx = {h: i for i, h in enumerate(hdr)}
    tot = {}
    for d in data:
        if d[idx["Metric Name"]] != "gpu__time_duration.sum":
            continue
        name = d[idx["Kernel Name"]].split("(")[0].replace("void ", "")
        v = float(d[idx["Metric Value"]].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(d[idx["Metric Unit"]], 1.0)
        n, t = tot.get(name, (0, 0.0))
        tot[name] = (n + 1, t + v)
    total = sum(t for _, t in tot.values())
    print("| kernel | launches | total us | us/launch | share |\n|---|---|---|---|---|")
    for name, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{name[:90]}` | {n} | {t:.1f} | {t / n:.1f} | {t / total:.3f} |")


def main():
    if sys.argv[1] == "--launches":
        return launch_shares(sys.argv[2])
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    out = []
    traffic = {}
    for d in data:
        name = d[idx["Kernel Name"]]
        for pat, short in SHORT:
            if pat in name and short not in traffic and "dram__bytes_read.sum" in idx:
                rd = to_bytes(d[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_read.sum"]])
                wr = to_bytes(d[idx["dram__bytes_write.sum"]], units[idx["dram__bytes_write.sum"]])
                traffic[short] = {"dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_launch": rd + wr,
                                  "duration_us_under_ncu": d[idx["gpu__time_duration.sum"]]}
        out.append(f"\n### {name[:110]}\n")
        out.append("| metric | value | unit |\n|---|---|---|")
        for key, label in WANT:
            if key in idx:
                out.append(f"| {label} (`{key}`) | {d[idx[key]]} | {units[idx[key]]} |")
    text = "\n".join(out)
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(text + "\n")
    if len(sys.argv) > 4:  # ncu_summary.py raw.csv out.md traffic.json <frames per launch of the profiled command>
        import json

        frames = int(sys.argv[4])
        for v in traffic.values():
            v["frames_per_launch"] = frames
            v["dram_bytes_per_frame"] = v["dram_bytes_per_launch"] / frames
        json.dump({"source": sys.argv[1], "how": "ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum of one launch",
                   "kernels": traffic}, open(sys.argv[3], "w"), indent=1)
    print(text)


if __name__ == "__main__":
    main()
