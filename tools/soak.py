"""Soak: interleave HRNet forwards of several shapes, the pipelined host loop, ShiftNet and the scoring kernels for a
while and check that every result stays bit-identical to its first value (races and stale-buffer bugs show up as drift).
python tools/soak.py [seconds]"""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import highres_net_b200 as hb
from oracle import hrnet_oracle, shiftnet_oracle
dev = torch.device("cuda:0")
secs = float(sys.argv[1]) if len(sys.argv) > 1 else 20.0
net = hb.HRNet(hrnet_oracle.DEFAULT_NETWORK_CONFIG).eval(); net.load_state_dict(hrnet_oracle.make_params(0)); net = net.to(dev)
sn = hb.ShiftNet().eval(); sn.load_state_dict(shiftnet_oracle.make_params(0)); sn = sn.to(dev)
g = torch.Generator().manual_seed(5)
cases = []
for (b, l, s) in [(32, 16, 128), (3, 5, 40), (2, 32, 128), (1, 1, 16), (4, 8, 256), (7, 4, 33)]:
    lrs = torch.rand(b, l, s, s, generator=g)
    al = (torch.rand(b, l, generator=g) > 0.15).float(); al[:, 0] = 1
    cases.append((lrs.pin_memory(), al.pin_memory(), lrs.to(dev), al.to(dev)))
pairs = torch.rand(70, 2, 128, 128, generator=g).to(dev)
hr = torch.rand(8, 384, 384, generator=g).to(dev); hm = (torch.rand(8, 384, 384, generator=g) > 0.1).float().to(dev)
shift = (torch.rand(8, 2, generator=g) * 2 - 1).to(dev)
ref = {}
def check(key, t):
    t = t.detach().cpu().clone()
    if key not in ref: ref[key] = t
    elif not torch.equal(ref[key], t): raise SystemExit(f"DRIFT in {key}: max|d| = {(ref[key] - t).abs().max().item()}")
t0, rounds = time.time(), 0
side = torch.cuda.Stream(device=dev)
while time.time() - t0 < secs:
    for i, (hl, ha, dl, da) in enumerate(cases):
        check(("dev", i), net(dl, da))
        p = net.forward_host_submit(hl, ha, device=dev)
        q = net.forward_host_submit(cases[(i + 1) % len(cases)][0], cases[(i + 1) % len(cases)][1], device=dev)
        check(("host", i), net.forward_host_wait(p)); check(("host", (i + 1) % len(cases)), net.forward_host_wait(q))
        with torch.cuda.stream(side):
            check(("shiftnet",), sn(pairs))
        sr = net(cases[0][2][:8], cases[0][3][:8])[:, 0]
        moved = hb.lanczos_shift(sr[None], shift, p=5)[0]
        best, xy, tab = hb.shift_cPSNR_argmax(moved, hr, hm, clip_sr=True)
        check(("lanczos",), moved); check(("cpsnr",), tab); check(("xy",), xy)
    rounds += 1
torch.cuda.synchronize()
print(json.dumps({"soak_seconds": round(time.time() - t0, 1), "rounds": rounds, "checked_results": len(ref), "drift": False}))
