"""128 -> 128 convs as cluster pairs with multicast A rows ("mcast" knob) vs the plain launch: bit-exactness on several
shapes and C2 timing on the same box.  python tools/mcast_ab.py"""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import highres_net_b200 as hb
from oracle import hrnet_oracle
dev = torch.device("cuda:0")
net = hb.HRNet(hrnet_oracle.DEFAULT_NETWORK_CONFIG).eval()
net.load_state_dict(hrnet_oracle.make_params(0)); net = net.to(dev)
torch.manual_seed(0)
for (b, l, s) in [(1, 2, 8), (2, 3, 33), (2, 4, 128), (1, 5, 100), (3, 2, 1), (1, 16, 64), (2, 8, 256)]:
    lrs = torch.rand(b, l, s, s, device=dev); al = (torch.rand(b, l, device=dev) > 0.2).float(); al[:, 0] = 1
    outs = []
    for mc in (1, 0):
        for ctas in (0, 6):
            net.debug_set(dev, "mcast", mc); net.debug_set(dev, "max_ctas", ctas)
            outs.append(net(lrs, al).clone())
    torch.cuda.synchronize()
    print(json.dumps({"shape": [b, l, s], "mcast_eq_plain": bool(torch.equal(outs[0], outs[2])), "ctas_invariant": bool(torch.equal(outs[0], outs[1]))}), flush=True)
net.debug_set(dev, "max_ctas", 0)
tl = [torch.rand(32, 16, 128, 128, device=dev) for _ in range(3)]; ta = torch.ones(32, 16, device=dev)
def timed(n=40):
    for i in range(10): net(tl[i % 3], ta)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n): net(tl[i % 3], ta)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for mc in (1, 0, 1, 0):
    net.debug_set(dev, "mcast", mc)
    ms = timed()
    time.sleep(1.5)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(3): net(tl[i % 3], ta)
    e1.record(); torch.cuda.synchronize()
    burst = e0.elapsed_time(e1) / 3
    net.profile_begin(dev)
    for _ in range(3): net(tl[0], ta)
    p = net.profile_end(dev)
    print(json.dumps({"mcast": mc, "ms": round(ms, 3), "burst_ms": round(burst, 3), "conv128": round(p["conv3x3_umma<128>"]["ms"] / 3, 3)}), flush=True)
