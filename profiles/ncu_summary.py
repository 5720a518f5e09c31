"""One line per captured kernel from an .ncu-rep: duration, DRAM bytes, tensor / L2 / DRAM / L1 utilisation.
usage: python profiles/ncu_summary.py X.ncu-rep"""
import csv
import io
import subprocess
import sys

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
want = {
    "gpu__time_duration.sum": "dur",
    "dram__bytes_read.sum": "dram_rd",
    "dram__bytes_write.sum": "dram_wr",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram%",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2%",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed": "l1%",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor%",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm%",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "occ%",
    "launch__registers_per_thread": "regs",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue%",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active": "lsu%",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed": "l1wave%",
}
idx = {h: i for i, h in enumerate(hdr)}
ik = idx["Kernel Name"]
for r in rows[2:]:
    parts = [r[ik][:60]]
    for k, lab in want.items():
        if k in idx:
            parts.append(f"{lab}={r[idx[k]]}{units[idx[k]] if lab in ('dur','dram_rd','dram_wr') else ''}")
    print("  ".join(parts))
