"""shift_cPSNR_argmax timing on one box: the 49-site window kernel (default) against the general shift-window kernel
("cpsnr_generic" knob), 32 and 512 imagesets of 384^2, plus the Lanczos shift at both sizes.  python tools/cpsnr_ab.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import highres_net_b200 as hb
dev = torch.device("cuda:0")


import time
import pynvml
pynvml.nvmlInit()
_nv = pynvml.nvmlDeviceGetHandleByIndex(0)


def sm_mhz():
    return pynvml.nvmlDeviceGetClockInfo(_nv, pynvml.NVML_CLOCK_SM)


def timed(fn, reps=5, inner=20):
    t0 = time.time()
    while time.time() - t0 < 1.5:          # clock ramp: the board idles at a few hundred MHz
        for _ in range(10):
            fn()
        torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(reps):
        e0.record()
        for _ in range(inner):
            fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / inner)
    return best


for n in (32, 512):
    sr = torch.rand(n, 384, 384, device=dev)
    hr = torch.rand(n, 384, 384, device=dev)
    hm = (torch.rand(n, 384, 384, device=dev) > 0.1).float()
    row = {"imagesets": n}
    for generic in (0, 1, 0, 1):
        hb.scoring_debug_set("cpsnr_generic", generic)
        ms = timed(lambda: hb.shift_cPSNR_argmax(sr, hr, hm))
        row["sm_mhz"] = sm_mhz()
        key = "generic" if generic else "window"
        row[key + "_ms"] = min(ms, row.get(key + "_ms", 1e9))
    hb.scoring_debug_set("cpsnr_generic", 0)
    for k in ("window", "generic"):
        row[k + "_GBps_alg"] = n * 1769472 / row[k + "_ms"] / 1e6
    shift = torch.rand(n, 2, device=dev) * 2 - 1
    ms = timed(lambda: hb.lanczos_shift(sr[None], shift, p=5, a=3, N=7))
    row["lanczos_ms"] = ms
    row["lanczos_GBps"] = n * 1179648 / ms / 1e6
    print(json.dumps(row), flush=True)
