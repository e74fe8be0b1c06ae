"""Forward time of the small configurations (launch-latency bound) for one or more builds: C1 = B2 L4 128x128,
B1 L16 128x128.  python tools/small_time.py [lib.so ...]"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, json
sys.path.insert(0, %r)
import torch
import highres_net_b200 as hb
from oracle import hrnet_oracle
dev = torch.device("cuda:0")
net = hb.HRNet(hrnet_oracle.DEFAULT_NETWORK_CONFIG).eval()
net.load_state_dict(hrnet_oracle.make_params(0)); net = net.to(dev)
out = {}
for name, (b, l) in {"C1_b2_l4": (2, 4), "b1_l16": (1, 16), "b4_l16": (4, 16)}.items():
    lrs = torch.rand(b, l, 128, 128, device=dev); al = torch.ones(b, l, device=dev)
    for _ in range(20): net(lrs, al)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for rep in range(5):
        e0.record()
        for _ in range(100): net(lrs, al)
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 100)
    out[name] = round(best, 4)
print(json.dumps(out))
''' % ROOT
libs = sys.argv[1:] or [None]
for rnd in range(2):
    for lib in libs:
        env = dict(os.environ)
        if lib: env["HRN_B200_LIB"] = os.path.abspath(lib)
        r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
        print(os.path.basename(lib or "in-tree"), rnd, r.stdout.strip() or r.stderr[-400:], flush=True)
