"""Same-box A/B timing of two builds of libhrn_b200.so (box-to-box variance is +-8 %, so only same-box
comparisons mean anything):  python tools/ab.py path/to/A.so path/to/B.so [rounds]"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, json
sys.path.insert(0, %r)
import torch
import highres_net_b200 as hb
from oracle import hrnet_oracle
dev = torch.device("cuda:0")
net = hb.HRNet(hrnet_oracle.DEFAULT_NETWORK_CONFIG).eval()
net.load_state_dict(hrnet_oracle.make_params(0)); net = net.to(dev)
lrs = torch.rand(32, 16, 128, 128, device=dev); al = torch.ones(32, 16, device=dev)
for _ in range(15): net(lrs, al)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(60): net(lrs, al)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 60
import time
bursts = []
for _ in range(3):
    time.sleep(1.5)                      # let the clocks recover: the un-throttled (burst) regime
    e0.record()
    for _ in range(3): net(lrs, al)
    e1.record(); torch.cuda.synchronize()
    bursts.append(e0.elapsed_time(e1) / 3)
net.profile_begin(dev)
for _ in range(5): net(lrs, al)
p = net.profile_end(dev)
sr = torch.rand(512, 384, 384, device=dev); hr = torch.rand(512, 384, 384, device=dev); hm = (torch.rand(512, 384, 384, device=dev) > 0.1).float()
sh = torch.rand(512, 2, device=dev) * 2 - 1
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n
print(json.dumps({"ms_per_step": ms, "burst3_ms": min(bursts), "conv64": p["conv3x3_umma<64>"]["ms"] / 5, "resblock64": p["resblock64_umma"]["ms"] / 5, "conv128": p["conv3x3_umma<128>"]["ms"] / 5,
                  "conv_init": p["conv_init"]["ms"] / 5, "decoder": p["decoder"]["ms"] / 5,
                  "lanczos512": t(lambda: hb.lanczos_shift(sr[None], sh, p=5)), "cpsnr512": t(lambda: hb.shift_cPSNR_argmax(sr, hr, hm))}))
''' % ROOT
libs = sys.argv[1:3]
rounds = int(sys.argv[3]) if len(sys.argv) > 3 else 2
for r in range(rounds):
    for name, lib in zip("AB", libs):
        env = dict(os.environ, HRN_B200_LIB=os.path.abspath(lib))
        out = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
        line = out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-400:]
        print(name, r, line, flush=True)
