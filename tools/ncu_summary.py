"""Summarise an .ncu-rep (CPU, no GPU needed) into a small CSV for profiles/:  python tools/ncu_summary.py report.ncu-rep out.csv"""
import csv, io, subprocess, sys
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_op_dmma.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__issue_active.avg.pct",
        "smsp__issue_active.avg.pct", "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__maximum_warps_per_active_cycle_pct",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__cycles_active.avg", "sm__cycles_elapsed.max"]
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
with open(out, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["kernel", "metric", "value", "unit"])
    for r in data:
        name = r[hdr.index("Kernel Name")]
        for i, h in enumerate(hdr):
            if h in KEYS or "stall" in h and h.endswith(".pct") and "not_issued" not in h and float((r[i] or "0").replace(",", "")) >= 5.0:
                w.writerow([name, h, r[i], units[i]])
print(open(out).read()[:6000])
