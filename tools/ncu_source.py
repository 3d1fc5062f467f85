"""Instruction mix, stall reasons and the hottest SASS lines of one kernel from an ncu report (source page).
   python tools/ncu_source.py gpurun_out/x.ncu-rep [kernel-index]"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
starts.append(len(rows))
lo, hi = starts[which], starts[which + 1]
print(rows[lo][1][:120])
h = rows[lo + 1]
ix = {n: i for i, n in enumerate(h)}
body = [r for r in rows[lo + 2:hi] if len(r) >= len(h)]
stall_cols = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
by_op, samples, stall = collections.Counter(), collections.Counter(), collections.Counter()
for r in body:
    op = r[ix["Source"]].split()
    if not op:
        continue
    o = (op[1] if op[0].startswith("@") else op[0]).split(".")[0]
    by_op[o] += int(r[ix["Instructions Executed"]])
    samples[o] += int(r[ix["# Samples"]])
    for c in stall_cols:
        stall[c] += int(r[ix[c]])
tot_s = sum(samples.values())
print("total instructions", sum(by_op.values()), "samples", tot_s)
for o, n in by_op.most_common(12):
    print(f"  {o:10s} {n:12d}  samples {samples[o]:8d} ({100 * samples[o] / tot_s:.1f}%)")
print("stalls:", [(k, v, f"{100 * v / tot_s:.1f}%") for k, v in stall.most_common(8)])
for r in sorted(body, key=lambda r: -int(r[ix["# Samples"]]))[:int(sys.argv[3]) if len(sys.argv) > 3 else 16]:
    print(r[ix["# Samples"]], r[ix["Instructions Executed"]], r[ix["Source"]][:70], {c[6:]: r[ix[c]] for c in stall_cols if int(r[ix[c]]) > 200})
