import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import highres_net_b200 as hb
from oracle import hrnet_oracle
dev = torch.device("cuda:0")
net = hb.HRNet(hrnet_oracle.DEFAULT_NETWORK_CONFIG).eval()
net.load_state_dict(hrnet_oracle.make_params(0)); net = net.to(dev)
lrs = torch.rand(32, 16, 128, 128).pin_memory(); al = torch.ones(32, 16).pin_memory()
out = torch.empty(32, 1, 384, 384).pin_memory()
tl, ta = lrs.to(dev), al.to(dev)
def timed(fn, n=40):
    for _ in range(10): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
print("device-resident", timed(lambda: net(tl, ta)))
for ch in (1, 2, 4, 8, 1, 4):
    net.debug_set(dev, "host_chunks", ch)
    print("host_chunks", ch, timed(lambda: net.forward_host(lrs, al, out_host=out, device=dev)))
