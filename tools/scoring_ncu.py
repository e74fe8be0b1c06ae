"""Target for ncu: the scoring kernels on N imagesets of 384^2 -- Lanczos shift, then the shift_cPSNR search (window
kernel, and with "generic" as argv[2] also the general kernel).
    ncu --set full --clock-control none --import-source on --kernel-name regex:"lanczos|cpsnr" -o out python tools/scoring_ncu.py 512"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import highres_net_b200 as hb
dev = torch.device("cuda:0")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
sr = torch.rand(n, 384, 384, device=dev); hr = torch.rand(n, 384, 384, device=dev); hm = (torch.rand(n, 384, 384, device=dev) > 0.1).float()
shift = torch.rand(n, 2, device=dev) * 2 - 1
for _ in range(2):
    moved = hb.lanczos_shift(sr[None], shift, p=5, a=3, N=7)
for generic in ((0, 1) if "generic" in sys.argv else (0,)):
    hb.scoring_debug_set("cpsnr_generic", generic)
    for _ in range(2):
        hb.shift_cPSNR_argmax(moved[0], hr, hm, clip_sr=True)
torch.cuda.synchronize()
