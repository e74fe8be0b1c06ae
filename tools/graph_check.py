"""Eager launches vs CUDA-graph replay of the same forward (are inter-kernel gaps significant?)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import highres_net_b200 as hb
from oracle import hrnet_oracle
dev = torch.device("cuda:0")
net = hb.HRNet(hrnet_oracle.DEFAULT_NETWORK_CONFIG).eval()
net.load_state_dict(hrnet_oracle.make_params(0)); net = net.to(dev)
lrs = torch.rand(32, 16, 128, 128, device=dev); al = torch.ones(32, 16, device=dev)
def timed(fn, n=60):
    for _ in range(10): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
print("eager ms/step", timed(lambda: net(lrs, al)))
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    for _ in range(3): out = net(lrs, al)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g, stream=s):
    out = net(lrs, al)
torch.cuda.synchronize()
print("graph ms/step", timed(lambda: g.replay()))
print("eager again  ", timed(lambda: net(lrs, al)))
