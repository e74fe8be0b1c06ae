"""ShiftNet eval forward (hrn_shiftnet_forward): error against the fp32 oracle and throughput in pairs/s.
python tools/shiftnet_time.py"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import highres_net_b200 as hb
from oracle import shiftnet_oracle
dev = torch.device("cuda:0")
params = shiftnet_oracle.make_params(0)
net = hb.ShiftNet().eval(); net.load_state_dict(params, strict=True); net = net.to(dev)
x = shiftnet_oracle.make_pairs(16, 5)
ref = shiftnet_oracle.shiftnet_forward(params, x).numpy()
got = net(torch.from_numpy(x).to(dev)).cpu().numpy()
print(json.dumps({"pairs": 16, "theta_max_abs": float(np.abs(ref).max()), "max_abs_err": float(np.abs(got - ref).max()),
                  "max_abs_err_between_pairs": float(np.abs((got - got.mean(0)) - (ref - ref.mean(0))).max()),
                  "spread_between_pairs": float((ref - ref.mean(0)).std())}), flush=True)
FLOP_PER_PAIR = 2 * (18 * 64 * 16384 + 576 * 64 * 16384 + 2 * 576 * 64 * 4096 + 576 * 128 * 1024 + 1152 * 128 * 1024
                     + 2 * 1152 * 128 * 256 + 32768 * 1024 + 2048)
for n, grouped, pool in ((32, 1, 1), (128, 1, 1), (512, 1, 1), (512, 1, 0), (512, 0, 0)):
    net.debug_set(dev, "img_group", grouped)
    net.debug_set(dev, "fused_pool", pool)
    xs = torch.rand(n, 2, 128, 128, device=dev)
    t0 = time.time()
    while time.time() - t0 < 1.0:
        for _ in range(5): net(xs)
        torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(3):
        e0.record()
        for _ in range(10): net(xs)
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 10)
    print(json.dumps({"pairs": n, "img_group": grouped, "fused_pool": pool, "ms": round(best, 4), "pairs_per_s": round(n / best * 1e3, 1),
                      "model_tflops": round(FLOP_PER_PAIR * n / best / 1e9, 1)}), flush=True)
