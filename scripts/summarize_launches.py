"""ncu launch list (--metrics gpu__time_duration.sum --csv) -> per-kernel markdown table for profiles/.

    python scripts/summarize_launches.py gpurun_out/launches.csv STEPS "title" "command" > profiles/xxx.md
STEPS = training steps the capture covers (totals are divided by it)."""
import csv, collections, sys

src, steps, title, cmd = sys.argv[1], int(sys.argv[2]), sys.argv[3], sys.argv[4]
rows = list(csv.reader(open(src)))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
H = rows[hdr]
ki, vi = H.index("Kernel Name"), H.index("Metric Value")
agg = collections.defaultdict(list)
for r in rows[hdr + 1:]:
    if len(r) > vi:
        name = r[ki].split("(")[0].replace("void ", "").replace("pivp::", "")
        agg[name].append(float(r[vi].replace(",", "")) / 1000)
tot = sum(sum(v) for v in agg.values())
n = sum(len(v) for v in agg.values())
print("# %s\n" % title)
print("Command: `%s`" % cmd)
print("(per-launch times are cold-cache and serialised: compare SHARES, not absolutes)\n")
print("%d launches over %d step(s) = %d per step, sum of kernel time %.1f us per step\n" % (n, steps, n // steps, tot / steps))
print("| kernel | launches / step | total us / step | avg us | share |\n|---|---:|---:|---:|---:|")
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print("| %s | %.1f | %.1f | %.2f | %.1f%% |" % (k, len(v) / steps, sum(v) / steps, sum(v) / len(v), 100 * sum(v) / tot))
