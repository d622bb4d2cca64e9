"""Per-kernel summary of an `ncu -i X.ncu-rep --page raw --csv` export: the judged metrics of every captured launch, one block per launch.
usage: python tools/ncu_csv_summary.py raw.csv [out.txt]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores"]
stalls = [i for i, h in enumerate(hdr) if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]
out = open(sys.argv[2], "w") if len(sys.argv) > 2 else sys.stdout
out.write("# %s\n" % sys.argv[1])
for r in rows[2:]:
    out.write("\n")
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            out.write("%-95s %-12s %s\n" % (w, units[i], r[i]))
    top = sorted(((float(r[i] or 0), hdr[i].replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")) for i in stalls), reverse=True)[:6]
    out.write("%-95s %-12s %s\n" % ("top stall reasons (warps per issue-active cycle)", "", ", ".join("%s %.2f" % (n, v) for v, n in top)))
