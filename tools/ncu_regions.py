"""Bucket the warp-state samples of one kernel by SASS region (60-instruction buckets) with the dominant opcodes / stalls.
   python tools/ncu_regions.py report.ncu-rep [min_samples]"""
import collections, csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = rows[1]; ix = {n: i for i, n in enumerate(h)}
body = [r for r in rows[2:] if len(r) >= len(h)]
tot = sum(int(r[ix['# Samples']]) for r in body)
print('total samples', tot, 'instructions', sum(int(r[ix['Instructions Executed']]) for r in body))
stall_cols = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
thr = int(sys.argv[2]) if len(sys.argv) > 2 else 1500
def opn(r):
    t = r[ix['Source']].split()
    return (t[1] if t[0].startswith('@') else t[0]).split('.')[0]
for b in range(0, len(body), 60):
    seg = body[b:b + 60]
    s = sum(int(r[ix['# Samples']]) for r in seg); n = sum(int(r[ix['Instructions Executed']]) for r in seg)
    if s < thr: continue
    ops = collections.Counter(opn(r) for r in seg)
    st = collections.Counter()
    for r in seg:
        for c in stall_cols: st[c[6:]] += int(r[ix[c]])
    print(b, s, f"{100*s/tot:.1f}%", n, ops.most_common(4), st.most_common(3))
