"""profiles/<tag>_traffic.json from an .ncu-rep (no GPU): per kernel, DRAM bytes per frame, duration,
issue utilisation, warp instructions.  python tools/ncu_traffic.py <rep> <frames in capture> <out.json>"""
import csv
import io
import json
import re
import subprocess
import sys

rep, frames, out = sys.argv[1], int(sys.argv[2]), sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "usecond": 1e-3, "msecond": 1.0, "nsecond": 1e-6, "second": 1e3}
kernels = {}
for r in rows[2:]:
    d = dict(zip(hdr, r))
    u = dict(zip(hdr, units))
    name = re.sub(r"^void ", "", d["Kernel Name"]).split("<")[0].split("(")[0]
    if name in kernels:
        continue
    val = lambda k: float(d[k]) * scale.get(u[k], 1.0)
    rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
    kernels[name] = {"frames_in_capture": frames, "dram_read_bytes": rd, "dram_write_bytes": wr,
                     "dram_bytes_per_frame": (rd + wr) / frames, "duration_ms": val("gpu__time_duration.sum"),
                     "issue_active_pct": float(d["smsp__issue_active.avg.pct_of_peak_sustained_active"]),
                     "warp_instructions": float(d["smsp__inst_executed.sum"])}
json.dump({"source": f"ncu --set full --clock-control none, bench.py --frames {frames} (C2), {rep}; summary in profiles/r1_ncu_summary.txt",
           "kernels": kernels}, open(out, "w"), indent=1)
print(json.dumps({k: round(v["dram_bytes_per_frame"] / 1e6, 2) for k, v in kernels.items()}))
