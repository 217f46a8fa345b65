"""Per-source-line stall samples of one kernel from an ncu report with --import-source:
   ncu -i X.ncu-rep --page source --csv --print-source cuda,sass --kernel-name regex:NAME > f.csv; python scripts/ncu_lines2.py f.csv [top]"""
import csv
import sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
cur_file = None
agg = []
hdr = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = {n: i for i, n in enumerate(r)}; continue
    if hdr and r[0].isdigit():
        try:
            agg.append((int(r[hdr["# Samples"]] or 0), int(r[hdr["Instructions Executed"]] or 0), cur_file, int(r[0]), r[1].strip()[:110]))
        except ValueError:
            pass
tot = sum(a[0] for a in agg) or 1
print("total samples", tot, "instructions", sum(a[1] for a in agg))
for s, n, f, ln, src in sorted(agg, reverse=True)[:top]:
    print(f"{100.0 * s / tot:5.1f}%  inst {n:9d}  {f}:{ln}  {src}")
