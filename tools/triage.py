"""Perf triage: time the conv classes with parts of the kernel disabled (results are garbage)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import highres_net_b200 as hb
from oracle import hrnet_oracle
dev = torch.device("cuda:0")
net = hb.HRNet(hrnet_oracle.DEFAULT_NETWORK_CONFIG).eval()
net.load_state_dict(hrnet_oracle.make_params(0)); net = net.to(dev)
lrs = torch.rand(32, 16, 128, 128, device=dev); al = torch.ones(32, 16, device=dev)
for flags in [int(x) for x in (sys.argv[1:] or ["0", "1", "2", "3", "4", "8", "15"])]:
    net.debug_set(dev, "debug_flags", flags)
    for _ in range(3): net(lrs, al)
    net.profile_begin(dev)
    for _ in range(4): net(lrs, al)
    p = net.profile_end(dev)
    print(f"flags={flags:2d} conv64 {(p['conv3x3_umma<64>']['ms'] + p['resblock64_umma']['ms'])/4:.3f} ms  conv128 {p['conv3x3_umma<128>']['ms']/4:.3f} ms", flush=True)
