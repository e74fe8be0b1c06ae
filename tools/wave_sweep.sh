#!/bin/bash
# Sustained-regime sweep of the wavefront knobs through bench.py --lean (same box, back to back).
mkdir -p gpurun_out
i=0
for knobs in "" "--knob wave_ring_rows=12" "--knob wave_ring_rows=24" "--knob wave_publish_rows=2" "--knob wave_ring_rows=12 --knob wave_publish_rows=2" "--no-wave" ""; do
  i=$((i+1))
  timeout 300 python bench.py --lean --no-cpu-baseline $knobs > gpurun_out/sweep_$i.json 2>> gpurun_out/sweep.err
  python - "$knobs" gpurun_out/sweep_$i.json <<'PY'
import json, sys
d = json.load(open(sys.argv[2])); r = d["roofline"]["per_class"]
print("%-55s value %.1f  ms %.3f  e2e %.1f  clk %s  fuse %.3f" % (sys.argv[1] or "(default)", d["value"], d["ms_per_step"], d["e2e"]["value"], d["clocks"]["sm_mhz"],
      r["fuse_wave"]["ms_per_step"] + r["conv3x3_umma<128>"]["ms_per_step"]))
PY
done
