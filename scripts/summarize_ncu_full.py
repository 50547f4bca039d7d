"""`ncu --set full` raw page (ncu -i X.ncu-rep --page raw --csv) -> one markdown table row per captured launch.

    python scripts/summarize_ncu_full.py "title" "command" gpurun_out/r2_full_*.csv > profiles/r02_ncu_full.md
"""
import csv, sys

METRICS = [
    ("gpu__time_duration.sum", "dur us", 1e-3, "ns"),
    ("dram__bytes_read.sum", "DRAM rd MB", None, "bytes"),
    ("dram__bytes_write.sum", "DRAM wr MB", None, "bytes"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %", 1, None),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %", 1, None),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1 %", 1, None),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %", 1, None),
    ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor (hmma) % of active", 1, None),
    ("sm__inst_issued.avg.pct_of_peak_sustained_active", "issue slots %", 1, None),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %", 1, None),
    ("launch__registers_per_thread", "regs", 1, None),
    ("launch__shared_mem_per_block_dynamic", "dyn smem KB", None, "smem"),
    ("launch__grid_size", "grid", 1, None),
    ("launch__block_size", "block", 1, None),
]
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1.0, "nsecond": 1.0, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6}

title, cmd, files = sys.argv[1], sys.argv[2], sys.argv[3:]
print("# %s\n" % title)
print("Command: `%s`\n" % cmd)
print("One row per captured launch (cold caches, serialised, ~40 replays per launch: read the RATIOS, not the absolute durations).\n")
print("| kernel | " + " | ".join(m[1] for m in METRICS) + " |")
print("|---|" + "---:|" * len(METRICS))
for f in files:
    rows = list(csv.reader(open(f)))
    try:
        hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    except StopIteration:
        print("| (%s: no launches captured) |" % f)
        continue
    H, units = rows[hdr], rows[hdr + 1]
    col = {n: i for i, n in enumerate(H)}
    for r in rows[hdr + 2:]:
        if len(r) <= col["Kernel Name"]:
            continue
        name = r[col["Kernel Name"]].split("(")[0].replace("void ", "").replace("pivp::", "")
        cells = []
        for key, _, sc, kind in METRICS:
            if key not in col or r[col[key]] in ("", "n/a"):
                cells.append("-")
                continue
            v = float(r[col[key]].replace(",", ""))
            u = units[col[key]].split("/")[0]
            if kind == "ns":
                cells.append("%.2f" % (v * SCALE.get(u, 1.0) / 1e3))
            elif kind in ("bytes", "smem"):
                b = v * SCALE.get(u, 1.0)
                cells.append("%.2f" % (b / 1e6) if kind == "bytes" else "%.1f" % (b / 1e3))
            else:
                cells.append("%.1f" % v if v != int(v) else "%d" % v)
        print("| %s | " % name + " | ".join(cells) + " |")
