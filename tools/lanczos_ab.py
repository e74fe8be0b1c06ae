"""Same-box A/B of the two N = 7 Lanczos kernels (register-window kernel vs TMA-fed tiles) on 512 x 384^2 and 32 x 384^2,
with a bit-exactness check between them.   python tools/lanczos_ab.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import highres_net_b200 as hb
dev = torch.device("cuda:0")
def timed(fn, n):
    for _ in range(20): fn()
    torch.cuda.synchronize()
    best = 1e9
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(5):
        e0.record()
        for _ in range(n): fn()
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / n)
    return best
for n, s in ((512, 384), (32, 384), (64, 132), (7, 96), (3, 8), (5, 260)):
    big = torch.rand(1, n, s, s, device=dev); sh = torch.rand(n, 2, device=dev) * 2 - 1
    outs = {}
    for name, knob in (("register window", 1), ("tma tiles", 0), ("register window", 1), ("tma tiles", 0)):
        hb.scoring_debug_set("lanczos_scalar", knob)
        outs[name] = hb.lanczos_shift(big, sh, p=5)
        ms = timed(lambda: hb.lanczos_shift(big, sh, p=5), 100)
        print(json.dumps({"images": n, "size": s, "kernel": name, "ms": round(ms, 4), "GBps": round(n * 2 * s * s * 4 / ms / 1e6, 1)}), flush=True)
    print("   bit-identical:", bool(torch.equal(outs["register window"], outs["tma tiles"])), flush=True)
hb.scoring_debug_set("lanczos_scalar", 0)
# the copy roofline at the same footprint: torch copy_ of the same 512 x 384^2 fp32 tensor (read + write, like MEASURED_PEAKS.json)
for n in (512, 2048):
    big = torch.rand(1, n, 384, 384, device=dev); dst = torch.empty_like(big)
    ms = timed(lambda: dst.copy_(big), 100)
    print(json.dumps({"images": n, "size": 384, "kernel": "torch copy_", "ms": round(ms, 4), "GBps": round(n * 2 * 384 * 384 * 4 / ms / 1e6, 1)}), flush=True)
