"""Times shift_cPSNR_argmax (four launches) on 32 and 512 imagesets of 384^2 for one or more builds of the library:
python tools/cpsnr_time.py [lib.so ...]   (same box, interleaved, 3 rounds)"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, json
sys.path.insert(0, %r)
import torch
import highres_net_b200 as hb
dev = torch.device("cuda:0")
out = {}
for n in (32, 512):
    sr = torch.rand(n, 384, 384, device=dev); hr = torch.rand(n, 384, 384, device=dev); hm = (torch.rand(n, 384, 384, device=dev) > 0.1).float()
    for _ in range(10): hb.shift_cPSNR_argmax(sr, hr, hm)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for rep in range(5):
        e0.record()
        for _ in range(20): hb.shift_cPSNR_argmax(sr, hr, hm)
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 20)
    out["ms_%%d" %% n] = best
    out["GBps_alg_%%d" %% n] = n * 1769472 / best / 1e6
print(json.dumps(out))
''' % ROOT
libs = sys.argv[1:] or [None]
for rnd in range(3):
    for lib in libs:
        env = dict(os.environ)
        if lib: env["HRN_B200_LIB"] = os.path.abspath(lib)
        r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
        print(os.path.basename(lib or "in-tree"), rnd, r.stdout.strip() or r.stderr[-400:], flush=True)
