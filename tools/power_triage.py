"""Where does the power go?  Sustained C2 forward loops with parts of the conv kernels disabled (results are then
garbage), nvidia-smi power / SM clock sampled meanwhile."""
import os, subprocess, sys, time, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import highres_net_b200 as hb
from oracle import hrnet_oracle
dev = torch.device("cuda:0")
net = hb.HRNet(hrnet_oracle.DEFAULT_NETWORK_CONFIG).eval()
net.load_state_dict(hrnet_oracle.make_params(0)); net = net.to(dev)
lrs = torch.rand(32, 16, 128, 128, device=dev); al = torch.ones(32, 16, device=dev)
names = {0: "full", 2: "no stores", 8: "no residual loads", 10: "no stores, no residual loads", 1: "no TMEM loads",
         4: "no TMA loads", 15: "MMA + barriers only", 31: "MMA only (no waits)"}
for flags in (0, 10, 4, 15, 31, 0):
    net.debug_set(dev, "debug_flags", flags)
    for _ in range(20): net(lrs, al)
    torch.cuda.synchronize()
    p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-lms", "50"],
                         stdout=subprocess.PIPE, text=True)
    time.sleep(0.3)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time(); n = 0
    e0.record()
    while time.time() - t0 < 2.5:
        for _ in range(10): net(lrs, al)
        torch.cuda.synchronize(); n += 10
    e1.record(); torch.cuda.synchronize()
    p.terminate(); out = p.communicate()[0]
    rows = [l.split(",") for l in out.strip().splitlines()][3:]
    clk = [float(r[0]) for r in rows if len(r) == 2]; pw = [float(r[1]) for r in rows if len(r) == 2]
    print(f"flags={flags:2d} {names[flags]:30s} ms/step {e0.elapsed_time(e1)/n:6.3f}  SM clock median {statistics.median(clk):6.0f} MHz  power median {statistics.median(pw):6.0f} W", flush=True)
