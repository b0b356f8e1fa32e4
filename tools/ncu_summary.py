"""Condense `ncu -i X.ncu-rep --page raw --csv` into the handful of metrics the roofline discussion uses."""
import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_static", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__grid_size",
        "launch__waves_per_multiprocessor", "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__t_sector_hit_rate.pct"]
idx = {h: i for i, h in enumerate(hdr)}
for r in rows[2:]:
    print("== %s  grid %s block %s" % (r[idx["Kernel Name"]], r[idx.get("Grid Size", 0)], r[idx.get("Block Size", 0)]))
    for w in want:
        if w in idx:
            print("   %-62s %s %s" % (w, r[idx[w]], units[idx[w]]))
