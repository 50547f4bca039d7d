"""Summarise an `ncu --set full` capture of the tcgen05 ConvLSTM kernel into profiles/<name>.json (default r02_ncu_full_halo_traffic.json).

    ncu -i gpurun_out/halo_full.ncu-rep --page raw --csv > gpurun_out/halo_full_raw.csv
    python scripts/ncu_traffic_summary.py gpurun_out/halo_full_raw.csv [r02_ncu_full_halo_traffic.json]

Writes, per launch (mean over the captured launches of one training step): DRAM bytes read + written, duration, tensor-pipe
utilisation -- bench.py reports `dram_bytes_per_launch` as roofline.traffic next to the algorithmic bytes."""
import csv, json, os, sys

src = sys.argv[1]
rows = list(csv.reader(open(src)))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
H, units = rows[hdr], rows[hdr + 1]
col = {n: i for i, n in enumerate(H)}
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3}


def val(r, name):
    i = col[name]
    return float(r[i].replace(",", "")) * scale.get(units[i].split("/")[0], 1.0)


launches = [r for r in rows[hdr + 2:] if len(r) > col["Kernel Name"] and "conv5x5_halo_tc_kernel" in r[col["Kernel Name"]]]
rd = [val(r, "dram__bytes_read.sum") for r in launches]
wr = [val(r, "dram__bytes_write.sum") for r in launches]
dur = [val(r, "gpu__time_duration.sum") for r in launches]
out = {"source": "ncu --set full --clock-control none, one training step (CDNA 64x64, batch 32, T=10, bf16), kernel conv5x5_halo_tc_kernel",
       "launches": len(launches), "dram_bytes_read_per_launch": sum(rd) / len(rd), "dram_bytes_write_per_launch": sum(wr) / len(wr),
       "dram_bytes_per_launch": (sum(rd) + sum(wr)) / len(rd), "avg_duration_us_under_ncu": sum(dur) / len(dur)}
for extra in ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum"):
    if extra in col:
        out[extra] = sum(val(r, extra) for r in launches) / len(launches)
dst = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", sys.argv[2] if len(sys.argv) > 2 else "r02_ncu_full_halo_traffic.json")
json.dump(out, open(dst, "w"), indent=1)
print(json.dumps(out, indent=1))
