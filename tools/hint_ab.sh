#!/bin/bash
# Same-box alternation of the L2 eviction hints through the lean bench line: debug flag 2048 switches all of them off,
# 4096 only the evict_last hint on the ring-row TMA reads of the wavefront kernel.
#   gpurun --timeout 1500 -- 'bash tools/hint_ab.sh [rounds]'
rounds=${1:-3}
mkdir -p gpurun_out
rm -f gpurun_out/h_*_[0-9].json
timeout 600 python -m pytest tests -m gpu -q -x -k "wave or resblock or forward or golden or soak" 2>&1 | tail -2
timeout 200 python bench.py --lean --no-cpu-baseline > /dev/null 2>&1     # warm the box
for i in $(seq 1 $rounds); do
  timeout 200 python bench.py --lean --no-cpu-baseline --knob debug_flags=2048 > gpurun_out/h_off_$i.json 2>/dev/null
  timeout 200 python bench.py --lean --no-cpu-baseline --knob debug_flags=4096 > gpurun_out/h_noring_$i.json 2>/dev/null
  timeout 200 python bench.py --lean --no-cpu-baseline > gpurun_out/h_on_$i.json 2>/dev/null
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/h_*_[0-9].json")):
    d = json.load(open(f)); r = d["roofline"]["per_class"]
    print(f.split("/")[-1], "value %.1f  ms %.3f  e2e %.1f  clk %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["clocks"]["sm_mhz"]),
          {k: round(v["ms_per_step"], 3) for k, v in r.items() if k in ("resblock64_umma", "fuse_wave")})
PY
