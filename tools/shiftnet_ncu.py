"""Target for ncu: two ShiftNet forwards on 512 pairs (the second one is the one to capture: 16 launches).
    ncu --set full --clock-control none --kernel-name regex:"umma|maxpool|center_planes|fc_finish" -s 16 -c 16 -o out python tools/shiftnet_ncu.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import highres_net_b200 as hb
from oracle import shiftnet_oracle
dev = torch.device("cuda:0")
net = hb.ShiftNet().eval(); net.load_state_dict(shiftnet_oracle.make_params(0)); net = net.to(dev)
x = torch.rand(512, 2, 128, 128, device=dev)
net(x); net(x)
torch.cuda.synchronize()
