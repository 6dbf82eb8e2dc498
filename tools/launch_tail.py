"""Prints the last N launches of an `ncu --metrics gpu__time_duration.sum --csv` launch list whose kernel name matches a
regex: short kernel name + template arguments, grid, microseconds.  Usage: launch_tail.py <csv> [regex] [N]"""
import csv, re, sys

path = sys.argv[1]
rx = re.compile(sys.argv[2] if len(sys.argv) > 2 else "tc_conv|last_dgrad")
n = int(sys.argv[3]) if len(sys.argv) > 3 else 13
rows = []
with open(path, newline="") as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum" or not rx.search(r["Kernel Name"]):
        continue
    name = r["Kernel Name"]
    m = re.search(r"(\w+)<([^>]*)>", name)
    short = "%s<%s>" % (m.group(1), m.group(2).replace("(int)", "").replace("(bool)", "")) if m else name.split("(")[0].split("::")[-1]
    rows.append((short, r["Grid Size"], float(r["Metric Value"].replace(",", "")) / 1e3))
tot = 0.0
for short, grid, us in rows[-n:]:
    tot += us
    print("%-60s %-14s %9.1f us" % (short, grid, us))
print("total %.1f us over %d launches" % (tot, min(n, len(rows))))
