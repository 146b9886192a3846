"""Digest of profiles/tower_trace.py output: per-item spans, inter-item gaps, stage periods, producer lead."""
import json
import sys

d = json.load(open(sys.argv[1]))
t = d["trace"]
t0 = t[0]
rel = lambda v: v - t0 if v else None
cyc, ns = t[1] - t[0], t[3] - t[2]
print("untraced us/launch", d["us_per_launch_untraced"], "traced cycles", cyc, "ns", ns, "GHz", cyc / ns)
st = [(rel(t[16 + 3 * i]), rel(t[16 + 3 * i + 1]), rel(t[16 + 3 * i + 2])) for i in range(600) if t[16 + 3 * i + 2]]
w = [rel(t[2000 + i]) for i in range(600) if t[2000 + i]]
act = [rel(t[3600 + i]) for i in range(400) if t[3600 + i]]
rows = [tuple(rel(t[2700 + 4 * j + k]) for k in range(4)) for j in range(200) if t[2700 + 4 * j + 3]]
# stages per channel half: 15 with the 4+1 row split (default build), 18 with 3+2; the stem item has one half
SPH = int(sys.argv[2]) if len(sys.argv) > 2 else 15
NS = 2 * SPH
bounds = [0, SPH]
while bounds[-1] + NS <= len(st):
    bounds.append(bounds[-1] + NS)
print("items traced", len(bounds) - 1)
tot_span = tot_gap = 0
for k in range(len(bounds) - 1):
    a, b = bounds[k], bounds[k + 1]
    span = st[b - 1][2] - st[a][0]
    gap = st[a][0] - (st[a - 1][2] if a else 0)
    waits = sum(s[1] - s[0] for s in st[a:b])
    issue = sum(s[2] - s[1] for s in st[a:b])
    if k:
        tot_span += span; tot_gap += gap
    print(f"item {k:2d} start {st[a][0]:7d} gap {gap:6d} span {span:6d} wait {waits:6d} issue {issue:6d}  first-stage wait {st[a][1]-st[a][0]}")
n = len(bounds) - 2
print("avg span", tot_span / n, "avg gap", tot_gap / n)
print("act issue times", act[:24])
print("weight producer lead over MMA (cycles between weight issue of stage i and MMA wait-done of stage i):")
lead = [st[i][1] - w[i] for i in range(min(len(st), len(w)))]
print("  min", min(lead), "avg", sum(lead) / len(lead), "max", max(lead))
print("stages of item 3 (wait, issue, period):")
a = bounds[3]
for i in range(a, a + NS):
    print("  ", i - a, st[i][1] - st[i][0], st[i][2] - st[i][1], st[i][2] - st[i - 1][2], "w_issue->ready", st[i][1] - w[i])
print("epilogue rows (wait, ld, work):")
for r in rows[10:30]:
    print("  ", r[0], r[1] - r[0], r[2] - r[1], r[3] - r[2])
