"""Aggregates an `ncu --page source --print-source cuda,sass --csv` dump per CUDA source line: share of executed warp
instructions, average active threads, top stall reason samples."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file = ""
hdr = None
agg = collections.defaultdict(lambda: [0, 0, 0])
text = {}
for r in rows:
    if len(r) >= 2 and r[0] in ("File Path", "File Name"):
        cur_file = r[1].split("/")[-1]
        continue
    if len(r) > 8 and r[0] == "Line No":
        hdr = r
        ie = hdr.index("Instructions Executed"); it = hdr.index("Thread Instructions Executed"); sm = hdr.index("# Samples")
        continue
    if hdr is None or len(r) <= ie:
        continue
    try:
        c = int(r[ie]); t = int(r[it]); s = int(r[sm])
    except ValueError:
        continue
    key = (cur_file, r[0])
    agg[key][0] += c; agg[key][1] += t; agg[key][2] += s
    if r[1].strip():
        text[key] = r[1].strip()[:110]
tot = sum(v[0] for v in agg.values()); tots = sum(v[2] for v in agg.values())
print("total warp instructions", tot, "samples", tots)
for key, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{v[0] / tot * 100:5.1f}% inst {v[2] / max(tots, 1) * 100:5.1f}% smp thr={v[1] / max(v[0], 1):4.1f} {key[0]}:{key[1]}  {text.get(key, '')}")
