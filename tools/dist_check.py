"""2-rank NCCL check of the sharded forward + score + gather (run under torchrun on a 2-GPU box):
   python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tools/dist_check.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import highres_net_b200 as hb
from highres_net_b200 import distributed as hd
from oracle import hrnet_oracle
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
net = hb.HRNet(hrnet_oracle.DEFAULT_NETWORK_CONFIG).eval(); net.load_state_dict(hrnet_oracle.make_params(0)); net = net.to(dev)
g = torch.Generator().manual_seed(3)
n = 7                                                       # ragged: 4 + 3 imagesets
lrs, alphas = torch.rand(n, 8, 64, 64, generator=g).to(dev), torch.ones(n, 8, device=dev)
full = net(lrs, alphas)                                     # every rank computes the whole batch as the check
hr = (torch.roll(full[:, 0], (1, -1), (1, 2)).clamp(0, 1) + 0.02).clamp(0, 1)
hm = torch.ones_like(hr)
sr, scores, xy = hd.sharded_forward_and_score(net, lrs, alphas, hr, hm)
ref_best, ref_xy, _ = hb.shift_cPSNR_argmax(full[:, 0], hr, hm, clip_sr=True)
ok = torch.equal(sr, full) and torch.allclose(scores, ref_best) and torch.equal(xy, ref_xy.to(torch.int64))
print(f"rank {rank}: sharded == full: {ok}; best shifts {xy.tolist()[:3]} ...", flush=True)
dist.destroy_process_group()
sys.exit(0 if ok else 1)
