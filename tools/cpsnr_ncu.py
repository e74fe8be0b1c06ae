"""One call of shift_cPSNR_argmax on 512 x 384^2 after a warm-up, for ncu -k regex:cpsnr_onepass_kernel -s 2 -c 1 ..."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import highres_net_b200 as hb
dev = torch.device("cuda:0")
n = 512
sr = torch.rand(n, 384, 384, device=dev); hr = torch.rand(n, 384, 384, device=dev); hm = (torch.rand(n, 384, 384, device=dev) > 0.1).float()
for _ in range(3):
    hb.shift_cPSNR_argmax(sr, hr, hm)
torch.cuda.synchronize()
