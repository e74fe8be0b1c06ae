"""Oracle (test infrastructure, not product): CPU restatement of the reference registration network in eval mode.

  * shiftnet_forward .... /root/reference/src/DeepNetworks/ShiftNet.py:49-75 (forward; layers :16-48)
  * register_batch ...... /root/reference/src/train.py:26-44

The arithmetic lives in PyTorch (conv2d, batch_norm with running statistics, max_pool2d, linear); this file restates the
graph with torch.nn.functional on seeded parameters.  Pinned against the unmodified reference module by
oracle/make_golden_shiftnet.py (tests/golden/shiftnet.npz)."""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

CIN = (2, 64, 64, 64, 64, 128, 128, 128)
COUT = (64, 64, 64, 64, 128, 128, 128, 128)
POOL = (False, True, False, True, False, True, False, False)


def make_params(seed: int = 0):
    """Seeded state_dict of ShiftNet(in_channel=1) with non-trivial BatchNorm statistics and a non-zero fc2 (the reference
    initialises fc2 to zero, ShiftNet.py:48, which would make every theta 0)."""
    g = torch.Generator().manual_seed(seed)

    def uni(shape, lo, hi):
        return torch.rand(shape, generator=g) * (hi - lo) + lo

    sd = {}
    for l in range(8):
        cin, cout = CIN[l], COUT[l]
        bound = (6.0 / (cin * 9)) ** 0.5
        sd[f"layer{l + 1}.0.weight"] = uni((cout, cin, 3, 3), -bound, bound)
        sd[f"layer{l + 1}.0.bias"] = uni((cout,), -0.1, 0.1)
        sd[f"layer{l + 1}.1.weight"] = uni((cout,), 0.5, 1.5)
        sd[f"layer{l + 1}.1.bias"] = uni((cout,), -0.2, 0.2)
        sd[f"layer{l + 1}.1.running_mean"] = uni((cout,), -0.2, 0.2)
        sd[f"layer{l + 1}.1.running_var"] = uni((cout,), 0.5, 1.5)
        sd[f"layer{l + 1}.1.num_batches_tracked"] = torch.tensor(7, dtype=torch.long)
    bound = (6.0 / 32768) ** 0.5
    sd["fc1.weight"] = uni((1024, 32768), -bound, bound)
    sd["fc1.bias"] = uni((1024,), -0.1, 0.1)
    sd["fc2.weight"] = uni((2, 1024), -0.05, 0.05)
    return sd


def make_pairs(n: int, seed: int = 0):
    """n pairs (reference, shifted + noisy view) of 128 x 128 crops in [0, 1], float32 (n, 2, 128, 128)."""
    rng = np.random.RandomState(7000 + seed)
    yy, xx = np.mgrid[0:128, 0:128].astype(np.float32)
    out = np.empty((n, 2, 128, 128), dtype=np.float32)
    for i in range(n):
        f = rng.uniform(0.05, 0.3, size=4)
        ref = 0.5 + 0.2 * np.sin(f[0] * xx + f[1] * yy) * np.cos(f[2] * yy - f[3] * xx) + 0.05 * rng.rand(128, 128)
        dy, dx = rng.randint(-2, 3, size=2)
        view = np.roll(ref, (dy, dx), (0, 1)) + 0.02 * rng.randn(128, 128) + rng.uniform(-0.05, 0.05)
        out[i, 0], out[i, 1] = ref, view
    return np.clip(out, 0, 1).astype(np.float32)


def shiftnet_forward(params, x, dtype=torch.float32):
    """ShiftNet.py:58-73 in eval mode.  x: (N, 2, 128, 128) array -> (N, 2) tensor."""
    x = torch.as_tensor(np.asarray(x)).to(dtype)
    p = {k: v.to(dtype) for k, v in params.items() if v.is_floating_point()}
    out = x - torch.mean(x, dim=(2, 3), keepdim=True)                                      # :58
    for l in range(8):
        k = f"layer{l + 1}"
        out = F.conv2d(out, p[k + ".0.weight"], p[k + ".0.bias"], padding=1)
        out = F.batch_norm(out, p[k + ".1.running_mean"], p[k + ".1.running_var"], p[k + ".1.weight"], p[k + ".1.bias"],
                           training=False, eps=1e-5)
        out = F.relu(out)
        if POOL[l]:
            out = F.max_pool2d(out, 2)
    out = out.reshape(-1, 128 * 16 * 16)                                                   # :67 (NCHW flatten)
    out = F.relu(F.linear(out, p["fc1.weight"], p["fc1.bias"]))                            # :70-71 (dropout: identity)
    return F.linear(out, p["fc2.weight"])                                                  # :72


def register_batch(params, lrs, reference):
    """train.py:26-44: lrs (B, V, H, W), reference (B, 1, H, W) -> thetas (B, V, 2)."""
    lrs, reference = np.asarray(lrs), np.asarray(reference)
    thetas = [shiftnet_forward(params, np.concatenate([reference, lrs[:, i:i + 1]], 1)) for i in range(lrs.shape[1])]
    return torch.stack(thetas, 1)
