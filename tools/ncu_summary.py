"""Summarise an .ncu-rep (read with `ncu -i ... --page raw --csv`) into a small table.

    python tools/ncu_summary.py gpurun_out/x.ncu-rep [out.csv]
"""
import csv
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem_wf%"),
    ("l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed", "lsu_wf%"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "bank_confl"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wf"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma%"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma_cyc%"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu_inst%"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
    ("launch__registers_per_thread", "regs"),
    ("launch__shared_mem_per_block_dynamic", "smem/blk"),
    ("launch__occupancy_limit_shared_mem", "lim_smem"),
    ("launch__occupancy_limit_registers", "lim_regs"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "st_barrier"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "st_short_sb"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "st_long_sb"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "st_mio"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "st_lg"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "st_math"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "st_wait"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "st_notsel"),
    ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "st_noinst"),
    ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "st_branch"),
    ("smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "st_disp"),
    ("smsp__inst_executed.sum", "inst"),
    ("lts__t_sector_hit_rate.pct", "l2hit%"),
]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    lines = []
    head = ["kernel", "grid", "block"] + [s for _, s in KEYS]
    lines.append(head)
    for r in rows[2:]:
        name = r[col["Kernel Name"]]
        name = name.replace("void ", "").replace("(LevelArgs<T1>)", "")
        line = [name[:90], r[col["Grid Size"]], r[col["Block Size"]]]
        for k, _ in KEYS:
            if k in col:
                v = r[col[k]]
                u = units[col[k]]
                try:
                    f = float(v.replace(",", ""))
                    v = f"{f:.4g}"
                except ValueError:
                    pass
                line.append(v + ("" if u in ("", "%", "ratio") or len(u) > 8 else " " + u))
            else:
                line.append("-")
        lines.append(line)
    if len(sys.argv) > 2:
        with open(sys.argv[2], "w", newline="") as f:
            csv.writer(f).writerows(lines)
    # transposed print
    for i, h in enumerate(head):
        print(f"{h:12s} " + " | ".join(f"{l[i][:34]:>34s}" if i == 0 else f"{l[i]:>34s}" for l in lines[1:]))


main()
