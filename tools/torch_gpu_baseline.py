"""The 'existing Blackwell kernel to beat' (SURVEY.md section 2a): the reference algorithm through PyTorch eager + cuDNN
on the same B200, (a) fp32 storage with TF32 convs (torch default, what the unmodified reference does on a GPU) and
(b) bf16 autocast + channels_last.  Uses the oracle's functional restatement (a tool, not the product)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import hrnet_oracle
import highres_net_b200 as hb
dev = torch.device("cuda:0")
params = {k: v.to(dev) for k, v in hrnet_oracle.make_params(0).items()}
b, l, s = 32, 16, 128
lrs = torch.rand(b, l, s, s, device=dev); al = torch.ones(b, l, device=dev)
def timed(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
res = {}
torch.backends.cudnn.benchmark = True
ms = timed(lambda: hrnet_oracle.hrnet_forward(params, lrs, al))
res["torch_eager_fp32_tf32"] = {"ms": ms, "imagesets_per_s": b / ms * 1e3}
def autocast_run():
    with torch.autocast("cuda", dtype=torch.bfloat16):
        return hrnet_oracle.hrnet_forward(params, lrs, al)
try:
    ms = timed(autocast_run)
    res["torch_eager_bf16_autocast"] = {"ms": ms, "imagesets_per_s": b / ms * 1e3}
except Exception as e:
    res["torch_eager_bf16_autocast"] = {"error": str(e)[:200]}
net = hb.HRNet(hrnet_oracle.DEFAULT_NETWORK_CONFIG).eval(); net.load_state_dict(hrnet_oracle.make_params(0)); net = net.to(dev)
ms = timed(lambda: net(lrs, al), 30)
res["highres_net_b200"] = {"ms": ms, "imagesets_per_s": b / ms * 1e3}
ref = hrnet_oracle.hrnet_forward(params, lrs[:2], al[:2]); out = net(lrs[:2], al[:2])
res["max_abs_diff_vs_torch_gpu_tf32"] = float((ref - out).abs().max())
print(json.dumps(res))
