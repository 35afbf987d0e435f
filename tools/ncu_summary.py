"""Summarise an .ncu-rep here (no GPU): per-kernel headline metrics + opcode mix + hottest source lines.
   python tools/ncu_summary.py gpurun_out/prof.ncu-rep [kernel-substring]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
flt = sys.argv[2] if len(sys.argv) > 2 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "smsp__inst_executed.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
        "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum"]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    if flt and flt not in d.get("Kernel Name", ""):
        continue
    print("=" * 100)
    for k in want:
        if k in d:
            print(f"  {k:70s} {d[k]}")
    for k in hdr:
        if "warp_issue_stalled" in k and k.endswith("per_warp_active.pct"):
            try:
                if float(d[k]) >= 3.0:
                    print(f"  stall {k.replace('smsp__warp_issue_stalled_', '').replace('_per_warp_active.pct', ''):40s} {d[k]}")
            except ValueError:
                pass
