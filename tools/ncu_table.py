"""Print the judged metrics of every kernel in an ncu raw-page CSV (ncu -i X.ncu-rep --page raw --csv) as a table."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
want = [("Kernel Name", "kernel"), ("gpu__time_duration.sum", "ms"), ("launch__grid_size", "grid"), ("launch__block_size", "blk"),
        ("launch__registers_per_thread", "regs"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("lts__t_sector_hit_rate.pct", "l2hit%"), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem%"),
        ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64%"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
        ("smsp__cycles_active.avg", "cycles"), ("sm__cycles_elapsed.max", "cyc_max")]
idx = [(hdr.index(k), n) for k, n in want if k in hdr]
print(" | ".join(n + ("[" + units[i] + "]" if units[i] else "") for i, n in idx))
for r in rows[2:]:
    print(" | ".join((r[i][:46] if n == "kernel" else r[i][:10]) for i, n in idx))
stalls = [i for i, h in enumerate(hdr) if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]
print("\nstall reasons (warps per issue-active cycle), top 5 per kernel:")
for r in rows[2:]:
    top = sorted(((float(r[i] or 0), hdr[i].replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")) for i in stalls), reverse=True)[:5]
    print("  %-40s %s" % (r[hdr.index("Kernel Name")][:40], ", ".join("%s %.2f" % (n, v) for v, n in top)))
