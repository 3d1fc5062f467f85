"""Warp-stall samples per CUDA source line of one kernel, from an ncu report with source counters:
   python tools/ncu_lines.py gpurun_out/x.ncu-rep [top-n]
Uses `ncu --page source --print-source cuda,sass`; the sources must be at the paths they were compiled from."""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
per_line = collections.defaultdict(lambda: collections.Counter())
text = {}
fpath, hdr = None, None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fpath = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r
        ix = {}
        for i, n in enumerate(hdr):
            ix.setdefault(n, i)
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    # rows: a CUDA line (Line No set, Address empty) followed by its SASS rows (Line No empty)
    if r[0]:
        cur = (fpath, int(r[0]))
        text[cur] = r[1].strip()
        continue
    try:
        n = int(r[ix["# Samples"]])
    except ValueError:
        continue
    c = per_line[cur]
    c["samples"] += n
    c["inst"] += int(r[ix["Instructions Executed"]] or 0)
    for name in hdr:
        if name.startswith("stall_") and "Not Issued" not in name:
            c[name[6:]] += int(r[ix[name]] or 0)
tot = sum(c["samples"] for c in per_line.values())
print("total samples", tot)
for k, c in sorted(per_line.items(), key=lambda kv: -kv[1]["samples"])[:top]:
    st = {n: v for n, v in c.items() if n not in ("samples", "inst") and v > 0.03 * c["samples"]}
    print(f"{100 * c['samples'] / tot:5.1f}%  {k[0]}:{k[1]:<4d} inst {c['inst']:>11d}  {text.get(k, '')[:90]}   {dict(sorted(st.items(), key=lambda kv: -kv[1])[:4])}")
