"""Summarise an ncu --set full report (.ncu-rep) into JSON: per kernel launch the metrics the roofline uses.
    python scripts/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x.json"""
import csv, json, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__inst_executed_pipe_fp64.sum", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum"]
res = []
for d in data:
    e = {}
    for i, h in enumerate(hdr):
        if h in want:
            v = d[i]
            try:
                v = float(v.replace(",", ""))
            except ValueError:
                pass
            e[h + (f" [{units[i]}]" if units[i] else "")] = v
    res.append(e)
print(json.dumps(res, indent=1))
