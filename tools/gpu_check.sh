#!/bin/bash
# One gpurun call: GPU tests, the full bench line, the ncu launch list and one `ncu --set full` forward step.
#   gpurun --timeout 1500 -- 'bash tools/gpu_check.sh [tag]'
tag=${1:-run}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${tag}_pytest.log
tail -5 gpurun_out/${tag}_pytest.log
timeout 600 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
head -c 1500 gpurun_out/${tag}_bench.json; echo
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_ncu_launches.csv \
    python bench.py --lean --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name 'regex:umma|median_anchor|live_lists|fuse_wave' -s 38 -c 19 \
    -f -o gpurun_out/${tag}_forward_full python bench.py --lean --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out | tail -8
