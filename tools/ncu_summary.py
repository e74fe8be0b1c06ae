"""Condense an `ncu --set full` report into the per-launch summary bench.py reads:
    ncu -i gpurun_out/<tag>_forward_full.ncu-rep --page raw --csv > raw.csv
    python tools/ncu_summary.py raw.csv profiles/r02_ncu_full_summary.json [--calls N --units M]
The summary records the source stamp of the library build that is in the tree NOW (libhrn_b200.so.stamp): take the capture
and run this script with the same build, so that bench.py can tell whether roofline.traffic is fresh."""
import csv, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FIELDS = {
    "t": "gpu__time_duration.sum", "rd": "dram__bytes_read.sum", "wr": "dram__bytes_write.sum",
    "tensor": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "dram_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "regs": "launch__registers_per_thread",
    "grid": "launch__grid_size", "smem": "launch__shared_mem_per_block_dynamic", "clk": "sm__cycles_elapsed.avg.per_second",
    "issue_active": "sm__inst_issued.avg.pct_of_peak_sustained_active", "fma_pipe": "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
}
TO_GB = {"Gbyte": 1.0, "Mbyte": 1e-3, "Kbyte": 1e-6, "byte": 1e-9}
TO_US = {"us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6, "usecond": 1.0, "msecond": 1e3, "nsecond": 1e-3, "second": 1e6}
args = sys.argv[1:]
opts = {}
while "--calls" in args or "--units" in args:
    for k in ("--calls", "--units"):
        if k in args:
            i = args.index(k); opts[k[2:]] = int(args[i + 1]); del args[i:i + 2]
rows = list(csv.reader(open(args[0])))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {k: hdr.index(v) for k, v in FIELDS.items() if v in hdr}
out = []
for r in data:
    name = r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "").replace("hrn::<", "").strip()
    e = {"name": name}
    for k, i in col.items():
        v = float(r[i].replace(",", "")) if r[i] not in ("", "n/a") else None
        if v is not None and k in ("rd", "wr"): v *= TO_GB.get(units[i], 1.0)
        if v is not None and k == "t": v *= TO_US.get(units[i], 1.0)
        e[k] = v
    out.append(e)
stamp_file = os.path.join(ROOT, "highres-net_b200", "csrc", "libhrn_b200.so.stamp")
doc = {"lib_stamp": open(stamp_file).read().strip() if os.path.exists(stamp_file) else None,
       "units": {"t": "us", "rd": "GB", "wr": "GB", "tensor": "% of elapsed", "clk": "Hz"},
       "dram_gb_total": sum((e["rd"] or 0) + (e["wr"] or 0) for e in out), "launches": out}
doc.update(opts)
json.dump(doc, open(args[1], "w"), indent=1)
for e in out:
    print(e["name"][:44].ljust(44), " ".join(f"{k}={e[k]:.3f}" for k in ("t", "rd", "wr", "tensor") if e.get(k) is not None))
print("total DRAM GB:", round(doc["dram_gb_total"], 3))
