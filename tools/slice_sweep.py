"""Does keeping one slice of the batch L2-resident between layers pay?  Sweeps the workspace cap (imagesets per slice) of
hrn_forward at C2 on one box (sustained loops, so every point sits under the same power cap)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import highres_net_b200 as hb
from oracle import hrnet_oracle
dev = torch.device("cuda:0")
net = hb.HRNet(hrnet_oracle.DEFAULT_NETWORK_CONFIG).eval()
net.load_state_dict(hrnet_oracle.make_params(0)); net = net.to(dev)
tl = [torch.rand(32, 16, 128, 128, device=dev) for _ in range(3)]; ta = torch.ones(32, 16, device=dev)
def timed(n=60):
    for i in range(15): net(tl[i % 3], ta)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n): net(tl[i % 3], ta)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for sets in (32, 1, 2, 4, 8, 16, 32, 1, 2, 32):
    net.debug_set(dev, "workspace_mb", 168 * sets + 8)
    print("imagesets/slice", sets, "ms/step %.3f" % timed(), flush=True)
