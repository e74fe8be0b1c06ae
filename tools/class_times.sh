#!/bin/bash
# Three lean bench lines with the per-class times of the sustained profile (quick look after a kernel change).
#   gpurun --timeout 900 -- 'bash tools/class_times.sh'
mkdir -p gpurun_out
for i in 1 2 3; do timeout 200 python bench.py --lean --no-cpu-baseline "$@" > gpurun_out/ct_$i.json 2>/dev/null; done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/ct_*.json")):
    d = json.load(open(f)); r = d["roofline"]["per_class"]
    print(f.split("/")[-1], "value %.1f  ms %.3f  clk %s" % (d["value"], d["ms_per_step"], d["clocks"]["sm_mhz"]),
          {k: round(v["ms_per_step"], 3) for k, v in r.items() if v["ms_per_step"] > 0.05})
PY
