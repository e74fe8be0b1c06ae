"""BASELINE.json configs beyond C2: correctness of one imageset vs the oracle and throughput (CUDA events)."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import highres_net_b200 as hb
from oracle import hrnet_oracle, scoring_oracle
dev = torch.device("cuda:0")
params = hrnet_oracle.make_params(0)
net = hb.HRNet(hrnet_oracle.DEFAULT_NETWORK_CONFIG).eval(); net.load_state_dict(params); net = net.to(dev)
def timed(fn, n):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for name, (b, l, s, iters) in {"C3 L32 128 (per-rank shard)": (32, 32, 128, 10), "C5 L8 512 B8": (8, 8, 512, 10), "C1 L4 128 B2": (2, 4, 128, 50)}.items():
    g = torch.Generator().manual_seed(11)
    lrs = torch.rand(b, l, s, s, generator=g); al = torch.ones(b, l)
    sr = net(lrs.to(dev), al.to(dev))
    t0 = time.time(); ref = hrnet_oracle.hrnet_forward(params, lrs[:1].numpy(), al[:1].numpy()).numpy(); t_cpu = time.time() - t0
    err = float(np.abs(sr[:1].cpu().numpy() - ref).max())
    ms = timed(lambda: net(lrs.to(dev) if False else tl, ta), iters) if False else None
    tl, ta = lrs.to(dev), al.to(dev)
    ms = timed(lambda: net(tl, ta), iters)
    fl = hrnet_oracle.flops_per_imageset(l, s, s) * b
    print(json.dumps({"config": name, "sr_max_err_vs_oracle": err, "ms": ms, "imagesets_per_s": b / ms * 1e3, "model_tflops": fl / ms / 1e9, "cpu_oracle_s_per_imageset": t_cpu}), flush=True)
# C4: full scoring path on 32 x 16-view imagesets
b, l, s = 32, 16, 128
g = torch.Generator().manual_seed(12)
tl, ta = torch.rand(b, l, s, s, generator=g).to(dev), torch.ones(b, l, device=dev)
sr = net(tl, ta)[:, 0]
shift = (torch.rand(b, 2, generator=g) * 2 - 1).to(dev)
hr = torch.roll(sr, (1, -2), (1, 2)).clamp(0, 1) + 0.02
hm = (torch.rand(b, 384, 384, generator=g) > 0.1).float().to(dev)
def score():
    moved = hb.lanczos_shift(sr[None], shift, p=5)[0]
    return hb.shift_cPSNR_argmax(moved, hr, hm, clip_sr=True)
ms_lz = timed(lambda: hb.lanczos_shift(sr[None], shift, p=5), 50)
moved = hb.lanczos_shift(sr[None], shift, p=5)[0]
ms_cp = timed(lambda: hb.shift_cPSNR_argmax(moved, hr, hm, clip_sr=True), 50)
ms_fw = timed(lambda: net(tl, ta), 10)
print(json.dumps({"config": "C4 scoring on 32 imagesets", "forward_ms": ms_fw, "lanczos_ms": ms_lz, "lanczos_GBps": 32 * 1179648 / ms_lz / 1e6,
                  "cpsnr_ms": ms_cp, "cpsnr_GBps_algorithmic": 32 * 1769472 / ms_cp / 1e6, "scoring_share_of_c4": (ms_lz + ms_cp) / (ms_fw + ms_lz + ms_cp)}), flush=True)
# larger scoring batches for a bandwidth number that is not launch-latency bound
big = torch.rand(1, 512, 384, 384, device=dev); sh = (torch.rand(512, 2, device=dev) * 2 - 1)
ms = timed(lambda: hb.lanczos_shift(big, sh, p=5), 20)
print(json.dumps({"config": "lanczos 512 x 384^2", "ms": ms, "GBps": 512 * 1179648 / ms / 1e6}), flush=True)
srb, hrb, hmb = torch.rand(512, 384, 384, device=dev), torch.rand(512, 384, 384, device=dev), (torch.rand(512, 384, 384, device=dev) > 0.1).float()
ms = timed(lambda: hb.shift_cPSNR_argmax(srb, hrb, hmb), 10)
print(json.dumps({"config": "cpsnr 512 x 384^2", "ms": ms, "GBps_algorithmic": 512 * 1769472 / ms / 1e6, "imagesets_per_s": 512 / ms * 1e3}), flush=True)
