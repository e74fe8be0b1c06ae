#!/bin/bash
# One gpurun call: GPU tests, the full bench line, the ncu launch list and one `ncu --set full` forward step.
#   gpurun --timeout 1800 -- 'bash tools/gpu_check.sh [tag] [extra ...]'
tag=${1:-run}
mkdir -p gpurun_out
python - <<'PY' || exit 9
import importlib.util, os, sys
spec = importlib.util.spec_from_file_location("b", "highres-net_b200/build.py"); m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
ok = os.path.exists(m.LIB) and os.path.exists(m.LIB + ".stamp") and open(m.LIB + ".stamp").read() == m._stamp()
print("library matches sources:", ok); sys.exit(0 if ok else 1)
PY
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${tag}_pytest.log
tail -8 gpurun_out/${tag}_pytest.log
if [[ " $* " == *" cpsnr "* ]]; then
  timeout 300 python tools/cpsnr_ab2.py > gpurun_out/${tag}_cpsnr_ab.log 2>&1; echo "cpsnr_ab rc=$?"; cat gpurun_out/${tag}_cpsnr_ab.log
  timeout 600 ncu --set full --clock-control none -k 'regex:lanczos7_tma|lanczos_shift7|cpsnr_onepass_kernel|cpsnr_window' -s 4 -c 2 -f -o gpurun_out/${tag}_scoring_full \
      python tools/scoring_ncu.py > gpurun_out/${tag}_ncu_scoring.log 2>&1; echo "ncu scoring rc=$?"
fi
if [[ " $* " == *" bench "* ]]; then
  timeout 600 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
  head -c 600 gpurun_out/${tag}_bench.json; echo; tail -3 gpurun_out/${tag}_bench.err
fi
if [[ " $* " == *" ab "* ]]; then
  for i in 1 2; do
    timeout 300 python bench.py --lean --no-cpu-baseline --no-wave > gpurun_out/${tag}_bench_nowave_$i.json 2>> gpurun_out/${tag}_bench.err; echo "no-wave $i rc=$?"
    timeout 300 python bench.py --lean --no-cpu-baseline > gpurun_out/${tag}_bench_wave_$i.json 2>> gpurun_out/${tag}_bench.err; echo "wave $i rc=$?"
  done
  python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/*_bench_*wave_*.json")):
    try:
        d = json.load(open(f)); r = d["roofline"]
        print(f.split("/")[-1], "value %.1f  ms %.3f  e2e %.1f  clk %s  fuse %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["clocks"]["sm_mhz"],
              {k: round(v["ms_per_step"], 3) for k, v in r["per_class"].items() if k in ("conv3x3_umma<128>", "fuse_wave", "gap")}))
    except Exception as e:
        print(f, "unreadable", e)
PY
fi
if [[ " $* " == *" ncu "* ]]; then
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_ncu_launches.csv \
      python bench.py --lean --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
  timeout 600 ncu --set full --clock-control none --import-source on --kernel-name 'regex:umma|median_anchor|live_lists|fuse_wave' -s ${NCU_SKIP:-22} -c ${NCU_COUNT:-11} \
      -f -o gpurun_out/${tag}_forward_full python bench.py --lean --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_full.log 2>&1; echo "ncu full rc=$?"
fi
ls -la gpurun_out | tail -8
