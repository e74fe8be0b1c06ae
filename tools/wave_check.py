"""First-light check of the wavefront kernels (encoder chain, fusion levels): SR bit-identical to the per-layer launches on
growing shapes.  Run under `timeout`: a protocol bug shows up as a trap after the bounded waits, never as a hang."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import highres_net_b200 as hb
from oracle import hrnet_oracle
dev = torch.device("cuda:0")
def make(**knobs):
    net = hb.HRNet(hrnet_oracle.DEFAULT_NETWORK_CONFIG).eval(); net.load_state_dict(hrnet_oracle.make_params(0)); net = net.to(dev)
    for k, v in knobs.items(): net.debug_set(dev, k, v)
    return net
ref_net = make(fuse_wave=0, enc_wave=0)
shapes = [(1, 2, 8), (1, 2, 16), (2, 3, 33), (1, 4, 64), (2, 4, 128), (1, 5, 100), (3, 2, 1), (1, 16, 64), (4, 8, 128), (32, 16, 128)]
if len(sys.argv) > 1: shapes = shapes[:int(sys.argv[1])]
for knobs in ({"fuse_wave": 0, "enc_wave": 1}, {"enc_wave": 1}, {"enc_wave": 1, "wave_streams": 1}, {"enc_wave": 1, "wave_streams": 3, "enc_ring_rows": 12}, {"wave_streams": 7, "wave_ring_rows": 8}):
    net = make(**knobs)
    for (b, l, s) in shapes:
        if "wave_streams" in knobs and b * l * s * s > 4 * 8 * 128 * 128: continue
        g = torch.Generator().manual_seed(b * 1000 + l * 10 + s)
        lrs = torch.rand(b, l, s, s, generator=g).to(dev); al = torch.ones(b, l, device=dev)
        if b > 1 and l > 2:
            al[1, l - 1:] = 0; lrs[1, l - 1:] = 0
            if l > 4: al[0, 2] = 0
        t0 = time.time()
        ref = ref_net(lrs, al); out = net(lrs, al); torch.cuda.synchronize()
        same = bool(torch.equal(ref, out))
        print(knobs, (b, l, s), "bit-identical" if same else f"DIFF max {float((ref - out).abs().max()):.3e}", f"{time.time() - t0:.2f}s", flush=True)
        if not same:
            d = (ref - out).abs()[:, 0]
            bad = d.amax(dim=(1, 2)); rows = d.amax(dim=2)
            print("  per-imageset max:", bad.cpu().numpy().round(5).tolist()[:8])
            r = rows[int(bad.argmax())].cpu().numpy(); print("  bad SR rows (of %d):" % r.shape[0], np.nonzero(r > 0)[0][:40].tolist())
lrs = torch.rand(32, 16, 128, 128, device=dev); al = torch.ones(32, 16, device=dev)
nets = [("per-layer launches", ref_net), ("fusion wavefront only", make()), ("encoder + fusion wavefronts", make(enc_wave=1))]
for name, m in nets + nets[:3]:
    for _ in range(20): m(lrs, al)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(100): m(lrs, al)
    e1.record(); torch.cuda.synchronize()
    m.profile_begin(dev)
    for _ in range(5): m(lrs, al)
    pr = m.profile_end(dev)
    enc = (pr["enc_wave"]["ms"] + pr["resblock64_umma"]["ms"] + pr["conv3x3_umma<64>"]["ms"]) / 5
    print(f"{name:30s} {e0.elapsed_time(e1) / 100:.3f} ms per C2 step   encoder convs {enc:.3f} ms", flush=True)
