"""ncu --csv (raw metrics page) -> a small JSON of per-launch counters for one kernel, stamped with the
sha256 of the kernel's sources so that bench.py can refuse stale numbers.

    python profiles/counters_to_json.py <ncu.csv> <out.json> <kernel regex> <source file> [<source file> ...]
"""
import csv
import hashlib
import json
import os
import re
import sys


def source_sha(paths):
    h = hashlib.sha256()
    for p in sorted(paths):
        with open(p, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def main():
    src, dst, pat = sys.argv[1], sys.argv[2], re.compile(sys.argv[3])
    files = sys.argv[4:]
    rows = []
    with open(src, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        rows.append(r)
    per_launch = {}
    for r in rows:
        if not pat.search(r.get("Kernel Name", "")):
            continue
        try:
            v = float(r["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        per_launch.setdefault(r["ID"], {"kernel": r["Kernel Name"], "grid": r.get("Grid Size"), "block": r.get("Block Size")})
        per_launch[r["ID"]][r["Metric Name"]] = {"value": v, "unit": r.get("Metric Unit", "")}
    launches = list(per_launch.values())
    if not launches:
        raise SystemExit("no launch matched " + sys.argv[3])
    last = launches[-1]                     # the warmest one
    out = {"kernel": last["kernel"], "grid": last["grid"], "block": last["block"], "n_launches_profiled": len(launches),
           "source_sha256": source_sha(files), "source_files": [os.path.relpath(p) for p in sorted(files)],
           "metrics": {k: v for k, v in last.items() if isinstance(v, dict)}}
    with open(dst, "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
