"""Per-kernel summary table of an ncu report (the metrics DESIGN.md / profiles/ quote).

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep
"""
import csv
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "time"),
    ("launch__grid_size", "grid"),
    ("launch__registers_per_thread", "regs"),
    ("smsp__inst_executed.sum", "warp-inst"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "lanes/inst"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"),
    ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "alu%"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma%"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64%"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu%"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu%"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit%"),
    ("lts__t_sector_hit_rate.pct", "L2 hit%"),
    ("dram__bytes_read.sum", "dram rd"),
    ("dram__bytes_write.sum", "dram wr"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "st long_sb"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "st wait"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "st not_sel"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "st math"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "st short_sb"),
    ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "st branch"),
    ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "st no_inst"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "st barrier"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "st lg"),
]


def main():
    out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    head, units, body = rows[0], rows[1], rows[2:]
    col = {n: i for i, n in enumerate(head)}
    names = [r[col["Kernel Name"]].replace("(StageParams)", "").replace("void ", "")[:22] for r in body]
    print("| metric | " + " | ".join(names) + " |")
    print("|---|" + "---|" * len(names))
    for key, label in WANT:
        if key not in col:
            continue
        i = col[key]
        vals = []
        for r in body:
            try:
                v = float(r[i].replace(",", ""))
                vals.append(f"{v:.3g}" if abs(v) < 1e6 else f"{v:.3e}")
            except ValueError:
                vals.append(r[i])
        print(f"| {label} ({units[i]}) | " + " | ".join(vals) + " |")


if __name__ == "__main__":
    main()
