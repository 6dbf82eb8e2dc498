"""Reads an .ncu-rep (ncu -i ... --page raw --csv) and prints one line per captured launch with the metrics the roofline
discussion uses.  Usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_ncu_summary.csv"""
import csv, io, subprocess, sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__cycles_active.avg", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_active.avg", "sm__cycles_elapsed.max",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed.sum", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"]
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
cols = {}
for w in WANT:
    for i, h in enumerate(hdr):
        if h == w:
            cols[w] = i
name_i = hdr.index("Kernel Name")
out = csv.writer(sys.stdout)
out.writerow(["kernel", "grid", "block"] + [w for w in WANT if w in cols] + ["dram_GBps"])
gi, bi = hdr.index("Grid Size"), hdr.index("Block Size")
for r in rows[2:]:
    if len(r) <= name_i:
        continue
    vals = [r[cols[w]] for w in WANT if w in cols]
    try:
        t = float(r[cols["gpu__time_duration.sum"]].replace(",", ""))
        by = float(r[cols["dram__bytes_read.sum"]].replace(",", "")) + float(r[cols["dram__bytes_write.sum"]].replace(",", ""))
        unit_t = rows[1][cols["gpu__time_duration.sum"]]
        unit_b = rows[1][cols["dram__bytes_read.sum"]]
        ts = t * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}.get(unit_t, 1e-9)
        bb = by * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit_b, 1.0)
        gbps = "%.1f" % (bb / ts / 1e9)
    except Exception:
        gbps = ""
    out.writerow([r[name_i][:90], r[gi], r[bi]] + vals + [gbps])
