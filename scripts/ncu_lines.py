"""Per-source-line instruction / stall-sample shares of one kernel from an .ncu-rep (ncu --set full --import-source on).
usage: python scripts/ncu_lines.py report.ncu-rep [top]"""
import csv, subprocess, sys
from collections import defaultdict

rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
d = defaultdict(lambda: [0, 0, ""]); cur = None; hdr = None
for r in rows:
    if len(r) >= 2 and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if len(r) > 2 and r[0] == "Line No": hdr = r; iS = hdr.index("# Samples"); iI = hdr.index("Instructions Executed"); continue
    if hdr and r and r[0].isdigit():
        try:
            e = d[(cur, int(r[0]))]; e[0] += int(r[iS]); e[1] += int(r[iI]); e[2] = r[1]
        except Exception:
            pass
ti = sum(v[1] for v in d.values()) or 1; ts = sum(v[0] for v in d.values()) or 1
print(f"total warp instructions {ti}, samples {ts}")
bf = defaultdict(lambda: [0, 0])
for (f, l), (s, i, _) in d.items(): bf[f][0] += s; bf[f][1] += i
for f, (s, i) in sorted(bf.items(), key=lambda kv: -kv[1][1]): print(f"{f:32s} {100*i/ti:5.1f}% inst {100*s/ts:5.1f}% smp")
for (f, l), (sm, i, src) in sorted(d.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{f[:16]:16s}{l:4d} {100*i/ti:5.1f}% inst {100*sm/ts:5.1f}% smp | {src.strip()[:100]}")
