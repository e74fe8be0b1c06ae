"""Sweep of the strip_split knob (row ranges per CTA, dealt round-robin) at C2 on one box."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import highres_net_b200 as hb
from oracle import hrnet_oracle
dev = torch.device("cuda:0")
net = hb.HRNet(hrnet_oracle.DEFAULT_NETWORK_CONFIG).eval()
net.load_state_dict(hrnet_oracle.make_params(0)); net = net.to(dev)
tl = [torch.rand(32, 16, 128, 128, device=dev) for _ in range(3)]; ta = torch.ones(32, 16, device=dev)
ref = None
def timed(n=40):
    for i in range(10): net(tl[i % 3], ta)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n): net(tl[i % 3], ta)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for k in (1, 4, 8, 16, 28, 56, 1, 8, 28):
    net.debug_set(dev, "strip_split", k)
    out = net(tl[0], ta)
    if ref is None: ref = out.clone()
    ms = timed()
    net.profile_begin(dev)
    for _ in range(3): net(tl[0], ta)
    p = net.profile_end(dev)
    print(json.dumps({"split": k, "ms": round(ms, 3), "same": bool(torch.equal(out, ref)), "conv64": round((p["conv3x3_umma<64>"]["ms"] + p["resblock64_umma"]["ms"]) / 3, 3), "conv128": round(p["conv3x3_umma<128>"]["ms"] / 3, 3)}), flush=True)
