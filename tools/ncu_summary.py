"""Condense an `ncu --set full` report into the per-launch summary bench.py reads (profiles/r01_ncu_full_summary.json):
    ncu -i gpurun_out/forward_full.ncu-rep --page raw --csv > raw.csv ; python tools/ncu_summary.py raw.csv out.json"""
import csv, json, sys
FIELDS = {
    "t": "gpu__time_duration.sum", "rd": "dram__bytes_read.sum", "wr": "dram__bytes_write.sum",
    "tensor": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "dram_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "regs": "launch__registers_per_thread",
    "grid": "launch__grid_size", "smem": "launch__shared_mem_per_block_dynamic", "clk": "sm__cycles_elapsed.avg.per_second",
}
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {k: hdr.index(v) for k, v in FIELDS.items() if v in hdr}
missing = [v for k, v in FIELDS.items() if k not in col]
if missing:
    print("metrics not in the report:", missing, file=sys.stderr)
out = []
for r in data:
    name = r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "").replace("hrn::<", "").strip()
    e = {"name": name}
    for k, i in col.items():
        e[k] = r[i]
    e["units"] = {k: units[i] for k, i in col.items()}
    out.append(e)
json.dump(out, open(sys.argv[2], "w"), indent=1)
for e in out:
    print(e["name"][:44].ljust(44), " ".join(f"{k}={e[k]}{e['units'][k]}" for k in ("t", "rd", "wr", "tensor") if k in e))
