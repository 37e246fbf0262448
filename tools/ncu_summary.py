"""Condense an `ncu --page raw --csv` dump into one line per profiled launch (the table committed under profiles/)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
cols = [("Kernel Name", "kernel", 18), ("gpu__time_duration.sum", "ms", 8),
        ("launch__registers_per_thread", "regs", 5), ("launch__occupancy_limit_shared_mem", "occS", 5),
        ("launch__occupancy_limit_registers", "occR", 5),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%", 6),
        ("smsp__inst_executed.sum", "winst", 11), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%", 7),
        ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64%", 6),
        ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma%", 6),
        ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu%", 6),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shwave", 11),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shconf", 11),
        ("l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed", "l1wave%", 7),
        ("dram__bytes_read.sum", "dramR", 9), ("dram__bytes_write.sum", "dramW", 9),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%", 6),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "st_bar", 6),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "st_lsb", 6),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "st_ssb", 6),
        ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "st_mio", 6),
        ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "st_math", 7),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "st_wait", 7),
        ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "st_lg", 6)]
idx = [(hdr.index(c) if c in hdr else -1, n, w) for c, n, w in cols]
print(" ".join(n.rjust(w) for _, n, w in idx))
for r in rows[2:]:
    out = []
    for i, n, w in idx:
        v = r[i] if i >= 0 else "-"
        if n == "kernel":
            v = v.split("(")[0].replace("bpc::", "")[:w]
        else:
            try:
                f = float(v.replace(",", ""))
                v = f"{f:.3g}" if abs(f) < 1e6 else f"{f:.3e}"
            except ValueError:
                pass
            if i >= 0 and n in ("dramR", "dramW"):
                v += units[i][:1]
        out.append(v.rjust(w))
    print(" ".join(out))
