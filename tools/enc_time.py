"""Timing triage of the encoder wavefront kernel at C2.  python tools/enc_time.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import highres_net_b200 as hb
from oracle import hrnet_oracle
dev = torch.device("cuda:0")
def make(**knobs):
    net = hb.HRNet(hrnet_oracle.DEFAULT_NETWORK_CONFIG).eval(); net.load_state_dict(hrnet_oracle.make_params(0)); net = net.to(dev)
    for k, v in knobs.items(): net.debug_set(dev, k, v)
    return net
lrs = torch.rand(32, 16, 128, 128, device=dev); al = torch.ones(32, 16, device=dev)
for name, kw in (("enc wave", dict(enc_wave=1)), ("no flag waits", dict(enc_wave=1, debug_flags=32)), ("no waits, no residual loads", dict(enc_wave=1, debug_flags=32 + 8)),
                 ("no waits, no ring stores", dict(enc_wave=1, debug_flags=32 + 2)), ("no waits, neither", dict(enc_wave=1, debug_flags=32 + 2 + 8))):
    net = make(**kw)
    for _ in range(10): net(lrs, al)
    net.debug_set(dev, "enc_stats", 1)
    for _ in range(10): net(lrs, al)
    print("==", name, "(10 forwards)", flush=True)
    net.debug_set(dev, "enc_stats", 0)
ref = make(fuse_wave=0, enc_wave=0)(lrs, al)
for name, kw in (("per-layer", dict(enc_wave=0)), ("enc wave", dict(enc_wave=1)), ("enc ring 16", dict(enc_wave=1, enc_ring_rows=16)),
                 ("enc ring 32", dict(enc_wave=1, enc_ring_rows=32)), ("per-layer", dict(enc_wave=0))):
    net = make(**kw)
    for _ in range(15): net(lrs, al)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(60): net(lrs, al)
    e1.record(); torch.cuda.synchronize()
    same = bool(torch.equal(net(lrs, al), ref))
    net.profile_begin(dev)
    for _ in range(5): net(lrs, al)
    pr = net.profile_end(dev)
    enc = (pr["enc_wave"]["ms"] + pr["resblock64_umma"]["ms"] + pr["conv3x3_umma<64>"]["ms"]) / 5
    print(f"{name:24s} {e0.elapsed_time(e1) / 60:7.3f} ms/step   encoder convs {enc:6.3f} ms   bit-identical: {same}", flush=True)
