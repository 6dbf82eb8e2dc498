"""Condenses `ncu -i <rep> --page source --csv` to the instructions that collect the most warp-stall samples, so the
report itself (hundreds of MB with source import) can stay on the GPU box.  Usage: ncu_source_top.py <rep> [top N]"""
import csv, io, subprocess, sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 45
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = raw.splitlines()
start = next((i for i, l in enumerate(lines) if l.startswith('"') and "Source" in l), 0)
rows = list(csv.reader(io.StringIO("\n".join(lines[start:]))))
if not rows:
    print("no source page; first lines:\n" + "\n".join(lines[:20]))
    sys.exit(0)
hdr = rows[0]
print("columns:", hdr)
def col(*names):
    for n in names:
        for i, h in enumerate(hdr):
            if h.strip().lower() == n.lower():
                return i
    for n in names:
        for i, h in enumerate(hdr):
            if n.lower() in h.lower():
                return i
    return None
c_src = col("Source")
c_smp = col("# Samples", "Warp Stall Sampling (All Samples)", "Samples")
c_exe = col("# Instructions Executed", "Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.lower().startswith("stall_") or h.lower().startswith("warp stall") and i != c_smp]
def num(s):
    try:
        return float(s.replace(",", ""))
    except Exception:
        return 0.0
body = [r for r in rows[1:] if len(r) == len(hdr)]
total = sum(num(r[c_smp]) for r in body) if c_smp is not None else 0
print("instructions: %d, total samples: %.0f" % (len(body), total))
for rank, r in enumerate(sorted(body, key=lambda r: -num(r[c_smp]))[:top]):
    st = sorted(((num(r[i]), hdr[i]) for i in stall_cols if num(r[i]) > 0), reverse=True)[:3]
    print("%5.1f%% %8.0f  exec %-10s %-90s %s" % (100 * num(r[c_smp]) / max(total, 1), num(r[c_smp]), r[c_exe] if c_exe is not None else "",
                                              r[c_src][:90], "; ".join("%s=%.0f" % (h, v) for v, h in st)))
