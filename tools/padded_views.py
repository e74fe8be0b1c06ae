"""Effect of dead-view skipping on Proba-V-like batches: config.json pads every imageset to n_views = 32, scenes hold
9..35 views (19 on average, paper).  Times B16 L32 128x128 dense vs skipped on the same box."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import highres_net_b200 as hb
from oracle import hrnet_oracle
dev = torch.device("cuda:0")
net = hb.HRNet(hrnet_oracle.DEFAULT_NETWORK_CONFIG).eval()
net.load_state_dict(hrnet_oracle.make_params(0)); net = net.to(dev)
B, L = 16, 32
rng = np.random.RandomState(0)
real = np.clip(np.round(rng.normal(19, 5, size=B)), 9, 32).astype(int)
lrs = torch.rand(B, L, 128, 128, device=dev); al = torch.ones(B, L, device=dev)
for i, n in enumerate(real):
    lrs[i, n:] = 0; al[i, n:] = 0
def timed(n=20):
    for _ in range(5): net(lrs, al)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): net(lrs, al)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
out = {"real_views": real.tolist(), "mean_real": float(real.mean())}
for skip in (1, 0, 1, 0):
    net.debug_set(dev, "skip_dead_views", skip)
    out.setdefault("ms_skip%d" % skip, []).append(timed())
out["speedup"] = min(out["ms_skip0"]) / min(out["ms_skip1"])
print(json.dumps(out))
