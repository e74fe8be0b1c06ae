"""Target for ncu: a few shift_cPSNR_argmax calls on 512 imagesets of 384^2 (window kernel, then the general one)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import highres_net_b200 as hb
dev = torch.device("cuda:0")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
sr = torch.rand(n, 384, 384, device=dev); hr = torch.rand(n, 384, 384, device=dev); hm = (torch.rand(n, 384, 384, device=dev) > 0.1).float()
for generic in (0, 1):
    hb.scoring_debug_set("cpsnr_generic", generic)
    for _ in range(2):
        hb.shift_cPSNR_argmax(sr, hr, hm)
torch.cuda.synchronize()
