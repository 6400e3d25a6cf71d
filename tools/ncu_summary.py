"""Compact per-kernel summary of an ncu report: python tools/ncu_summary.py report.ncu-rep [out.md]"""
import csv
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "time"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs"),
    ("launch__shared_mem_per_block_dynamic", "smem/blk"),
    ("launch__occupancy_limit_registers", "occ_lim_regs"), ("launch__occupancy_limit_shared_mem", "occ_lim_smem"),
    ("launch__occupancy_limit_warps", "occ_lim_warps"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved_occ%"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active%"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "pipe_fma%"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "pipe_alu%"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "pipe_xu%"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "pipe_lsu%"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_active", "l1tex%"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("dram__bytes_read.sum", "dram_read"), ("dram__bytes_write.sum", "dram_write"),
    ("lts__t_sector_hit_rate.pct", "l2_hit%"),
    ("smsp__inst_executed.sum", "warp_inst"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wavefronts"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_bank_conflicts"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall_barrier"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall_long_sb"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall_short_sb"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall_mio"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall_math"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall_wait"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall_not_sel"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall_lg"),
    ("smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "stall_dispatch"),
    ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall_no_inst"),
]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    seen = {}
    for d in data:
        name = d[idx["Kernel Name"]].split("(")[0]
        seen.setdefault(name, []).append(d)
    out = []
    for name, ds in seen.items():
        d = ds[-1]
        out.append(f"### {name}  ({len(ds)} captured launches; last shown)")
        out.append("")
        out.append("| metric | value | unit |")
        out.append("|---|---|---|")
        for key, label in KEYS:
            if key in idx:
                out.append(f"| {label} | {d[idx[key]]} | {units[idx[key]]} |")
        out.append("")
    text = "\n".join(out)
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(text + "\n")
    print(text)


if __name__ == "__main__":
    main()
