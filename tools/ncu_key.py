"""Key metrics of an ncu --set full report (development aid): python tools/ncu_key.py report.ncu-rep [more substrings]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct", "l1tex__throughput.avg.pct", "sm__throughput.avg.pct", "smsp__issue_active.avg.pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit",
        "smsp__average_warps_issue_stalled", "sm__inst_executed.avg.per_cycle_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "launch__grid_size", "launch__block_size",
        "smsp__warps_eligible.avg.per_cycle_active", "sm__ops_path_tensor_op_utchmma"] + sys.argv[2:]
for r in rows[2:]:
    print("==", r[hdr.index("Kernel Name")][:60] if "Kernel Name" in hdr else "")
    for h, u, v in zip(hdr, units, r):
        if any(w in h for w in want):
            print(f"  {h} [{u}] = {v}")
