"""One line per kernel launch of an `ncu --set full` report, the columns the roofline discussion needs.
    python scratch/ncu_full_summary.py gpurun_out/<rep>.ncu-rep [kernel regex] > profiles/<name>.csv"""
import csv, io, re, subprocess, sys
rep = sys.argv[1]
rx = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "sm__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem"]
cols = [h for h in want if h in hdr]
idx = [hdr.index(h) for h in cols]
w = csv.writer(sys.stdout)
w.writerow(cols); w.writerow([units[i] for i in idx])
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    if rx and not rx.search(name):
        continue
    r = list(r)
    r[hdr.index("Kernel Name")] = re.sub(r"\(.*", "", name.replace("b200surv::", "").replace("<unnamed>::", "").replace("unnamed>::", ""))
    w.writerow([r[i] for i in idx])
