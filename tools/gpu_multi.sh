#!/bin/bash
# N-GPU check on one box: the two-GPU tests, then the bench line under torchrun (NCCL gather of SR inside the timed region).
#   gpurun --gpus 2 --timeout 1500 -- 'bash tools/gpu_multi.sh 2 tag'
n=${1:-2}; tag=${2:-multi}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 600 python -m pytest tests -m gpu -q -k "nccl or two_devices" > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${tag}_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n \
    > gpurun_out/${tag}_bench_${n}gpu.json 2> gpurun_out/${tag}_bench_${n}gpu.err; echo "bench rc=$?"
python - gpurun_out/${tag}_bench_${n}gpu.json <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
print("n_gpus", d["n_gpus"], "value %.1f" % d["value"], "e2e %.1f" % d["e2e"]["value"], "ms %.3f" % d["ms_per_step"], d["config"]["collective"][:90])
print("c3 %.1f c5 %.1f c4 %.1f" % (d["configs"]["c3_shard"]["value"], d["configs"]["c5"]["value"], d["scoring"]["c4_imagesets_per_s"]), d["scoring"]["c4"]["collective"][:60])
PY
tail -5 gpurun_out/${tag}_bench_${n}gpu.err
