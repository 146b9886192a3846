"""Per-source-line dynamic instruction counts of one kernel from an ncu report taken with --import-source on:
ncu -i REP --page source --print-source cuda,sass --csv | python profiles/line_map.py [min_count]"""
import collections
import csv
import sys

rows = list(csv.reader(sys.stdin))
thr = int(sys.argv[1]) if len(sys.argv) > 1 else 150000
cur = line = None
per, mov, stall, src, tot = collections.Counter(), collections.Counter(), collections.Counter(), {}, 0
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r[0] in ("Line No", "Function Name"):
        continue
    if r[0] != "":
        try:
            line = int(r[0])
            src[(cur, line)] = r[1].strip()[:110]
        except ValueError:
            pass
        continue
    try:
        n = int(r[7] or 0)
        smp = int(r[4] or 0)
    except (ValueError, IndexError):
        continue
    per[(cur, line)] += n
    stall[(cur, line)] += smp
    tot += n
    if "MOV" in r[3]:
        mov[(cur, line)] += n
ts = sum(stall.values())
print(f"total warp instructions (inlined lines counted once per attribution) {tot}, stall samples {ts}")
for k in sorted(per):
    if per[k] >= thr:
        print(f"{k[0][:9]}:{k[1]:4d} {per[k]:8d} {100 * per[k] / tot:4.1f}%  mov {mov[k]:7d}  samples {100 * stall[k] / max(ts, 1):4.1f}% | {src.get(k)}")
