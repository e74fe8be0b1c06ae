"""Run the C2 forward in a loop for a few seconds while nvidia-smi samples SM clock and power."""
import os, subprocess, sys, time, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import highres_net_b200 as hb
from oracle import hrnet_oracle
dev = torch.device("cuda:0")
net = hb.HRNet(hrnet_oracle.DEFAULT_NETWORK_CONFIG).eval()
net.load_state_dict(hrnet_oracle.make_params(0)); net = net.to(dev)
lrs = torch.rand(32, 16, 128, 128, device=dev); al = torch.ones(32, 16, device=dev)
for _ in range(3): net(lrs, al)
torch.cuda.synchronize()
p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown",
                      "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE, text=True)
time.sleep(0.5)
t0 = time.time(); n = 0
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
while time.time() - t0 < 4.0:
    for _ in range(10): net(lrs, al)
    torch.cuda.synchronize(); n += 10
e1.record(); torch.cuda.synchronize()
time.sleep(0.2); p.terminate(); out = p.communicate()[0]
rows = [l.split(",") for l in out.strip().splitlines()]
clk = [float(r[0]) for r in rows if len(r) >= 2]; pw = [float(r[1]) for r in rows if len(r) >= 2]
print("forwards", n, "ms/forward", e0.elapsed_time(e1) / n)
print("samples", len(clk), "clock MHz min/median/max", min(clk), statistics.median(clk), max(clk), "power W median/max", statistics.median(pw), max(pw))
print("clock samples:", clk[::4])
print("reasons seen:", {tuple(x.strip() for x in r[2:]) for r in rows})
