"""Top warp-stall sites of the first kernel in an .ncu-rep (SASS view, with source line correlation when -lineinfo was
used):  python tools/ncu_top_stalls.py report.ncu-rep [n] [view]   (view: sass | cuda,sass)"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
view = sys.argv[3] if len(sys.argv) > 3 else "sass"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", view], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
his = [i for i, r in enumerate(rows) if r and r[0] in ("Address", "#")]
if not his:
    print(out[:2000])
    sys.exit(1)
hi = his[0]
hdr = rows[hi]
ends = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name" and i > hi]
body = rows[hi + 1:(ends[0] if ends else len(rows))]
si, src = hdr.index("# Samples"), hdr.index("Source")
tot = sum(int(r[si]) for r in body if len(r) > si and r[si].isdigit())
print(rows[0][1][:120] if rows[0] else "", "| total samples", tot, "| rows", len(body))
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
agg = {}
for r in body:
    for i in stall_cols:
        if len(r) > i and r[i].isdigit():
            agg[hdr[i]] = agg.get(hdr[i], 0) + int(r[i])
print("stall totals:", sorted(((v, k) for k, v in agg.items() if v), reverse=True)[:10])
for r in sorted(body, key=lambda r: -int(r[si]) if len(r) > si and r[si].isdigit() else 0)[:n]:
    st = sorted(((int(r[i]), hdr[i][6:]) for i in stall_cols if r[i].isdigit() and int(r[i]) > 0), reverse=True)[:3]
    print(r[si].rjust(7), r[0][-5:], r[src][:100].ljust(100), st)
