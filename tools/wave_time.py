"""Timing triage of the wavefront fusion kernel at C2: hand-over waits / release fences switched off (results garbage),
ring depth, stream count.  python tools/wave_time.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import highres_net_b200 as hb
from oracle import hrnet_oracle
dev = torch.device("cuda:0")
def make(wave, **knobs):
    net = hb.HRNet(hrnet_oracle.DEFAULT_NETWORK_CONFIG).eval(); net.load_state_dict(hrnet_oracle.make_params(0)); net = net.to(dev)
    net.debug_set(dev, "fuse_wave", wave)
    for k, v in knobs.items(): net.debug_set(dev, k, v)
    return net
lrs = torch.rand(32, 16, 128, 128, device=dev); al = torch.ones(32, 16, device=dev)
def timed(m, n=60):
    for _ in range(15): m(lrs, al)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): m(lrs, al)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
configs = [("three launches", dict(wave=0)), ("wave", dict(wave=1)), ("wave ring 12", dict(wave=1, wave_ring_rows=12)),
           ("wave ring 24", dict(wave=1, wave_ring_rows=24)), ("three launches", dict(wave=0)), ("wave", dict(wave=1))]
ref = make(0)(lrs, al)
for name, kw in (("wave", dict(wave=1)), ("wave no flag waits", dict(wave=1, debug_flags=32))):
    kw = dict(kw); wave = kw.pop("wave")
    net = make(wave, **kw)
    for _ in range(10): net(lrs, al)
    net.debug_set(dev, "wave_stats", 1)
    for _ in range(10): net(lrs, al)
    print("==", name, "(10 forwards = 40 level launches)", flush=True)
    net.debug_set(dev, "wave_stats", 0)

for name, kw in configs:
    kw = dict(kw); wave = kw.pop("wave")
    net = make(wave, **kw)
    ms = timed(net)
    outs = [net(lrs, al) for _ in range(30)]
    bad = [i for i, o in enumerate(outs) if not torch.equal(o, ref)]
    same = not bad
    if bad:
        d = (outs[bad[0]] - ref).abs()[:, 0]; per = d.amax(dim=(1, 2))
        print('   mismatching forwards', bad, 'imagesets', [int(i) for i in torch.nonzero(per > 0)[:, 0][:10]], 'max', float(d.max()), 'rows', (lambda rr: (int(rr.min()), int(rr.max()), int(rr.numel())))(torch.nonzero(d[int(per.argmax())].amax(dim=1) > 0)[:, 0]))
    prof = None
    net.profile_begin(dev)
    for _ in range(5): net(lrs, al)
    prof = net.profile_end(dev)
    fuse = (prof["fuse_wave"]["ms"] + prof["conv3x3_umma<128>"]["ms"]) / 5
    print(f"{name:28s} {ms:7.3f} ms/step   fusion stage {fuse:6.3f} ms   bit-identical x10: {same}", flush=True)
