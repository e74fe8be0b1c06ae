"""Longer same-box alternation of the "mcast" knob (sustained, power-capped regime): python tools/mcast_long.py [rounds]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import highres_net_b200 as hb
from oracle import hrnet_oracle
dev = torch.device("cuda:0")
net = hb.HRNet(hrnet_oracle.DEFAULT_NETWORK_CONFIG).eval()
net.load_state_dict(hrnet_oracle.make_params(0)); net = net.to(dev)
tl = [torch.rand(32, 16, 128, 128, device=dev) for _ in range(5)]; ta = torch.ones(32, 16, device=dev)
rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 6
for _ in range(100): net(tl[0], ta)
torch.cuda.synchronize()
res = {0: [], 1: []}
for r in range(rounds):
    for mc in (1, 0):
        net.debug_set(dev, "mcast", mc)
        for i in range(10): net(tl[i % 5], ta)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(100): net(tl[i % 5], ta)
        e1.record(); torch.cuda.synchronize()
        res[mc].append(round(e0.elapsed_time(e1) / 100, 3))
print(json.dumps({"mcast_ms": res[1], "plain_ms": res[0], "mean_mcast": sum(res[1]) / rounds, "mean_plain": sum(res[0]) / rounds}))
