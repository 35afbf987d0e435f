"""Attribute executed instructions of one kernel in an .ncu-rep to CUDA source lines (no GPU needed).
   python tools/ncu_lines.py <rep> <kernel-substring-in-ncu-name> <cubin> <mangled-substring> [top]
The SASS page of the report lists instructions in program order; nvdisasm -g lists the same function
with '//## File "...", line N' markers, so the two are joined by position."""
import csv
import io
import re
import subprocess
import sys

rep, kern, cubin = sys.argv[1], sys.argv[2], sys.argv[3]
mang = sys.argv[4]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
secs, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        secs.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
sec = [s for s in secs if kern in s["name"]][0]
hdr = sec["rows"][0]
iex, isrc, ist = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)")
counts = [(int(r[iex]), int(r[ist]), r[isrc]) for r in sec["rows"][1:] if len(r) > iex and r[iex].isdigit()]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
mangled = None
lines, cur_line, infn = [], None, False
for ln in dis:
    if ln.startswith(".text."):
        infn = mang in ln
        continue
    if not infn:
        continue
    m = re.search(r'//## File ".*?([^/"]+)", line (\d+)', ln)
    if m:
        cur_line = (m.group(1), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4}\*/", ln):
        lines.append(cur_line)
assert len(lines) == len(counts), (len(lines), len(counts))
agg, tot = {}, 0
for (n, st, _), l in zip(counts, lines):
    a = agg.setdefault(l, [0, 0])
    a[0] += n
    a[1] += st
    tot += n
srcs = {}
print(f"{sec['name'][:80]}: {tot} warp-instructions")
for l, (n, st) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    text = ""
    if l:
        try:
            if l[0] not in srcs:
                import glob
                srcs[l[0]] = open(glob.glob(f"timg_b200/csrc/{l[0]}")[0]).read().splitlines()
            text = srcs[l[0]][l[1] - 1].strip()[:100]
        except Exception:
            pass
    print(f"{100 * n / tot:5.1f}%  stall {st:6d}  {l}  {text}")
