"""Oracle (test infrastructure, not product): functional CPU restatement of
HRNet.forward in plain torch fp32/fp64 ops.

Follows /root/reference/src/DeepNetworks/HRNet.py:
  * median anchor + (view, anchor) pairing ........ HRNet.py:196-204
  * Encoder (conv+PReLU, residual blocks, conv) .... HRNet.py:51-74, 17-33
  * recursive alpha-masked pair fusion ............. HRNet.py:99-134
  * Decoder (stride-3 deconv, PReLU, 1x1 conv) ..... HRNet.py:147-169

The arithmetic itself (conv2d, conv_transpose2d, prelu, median) lives in
PyTorch (third party, not under /root/reference; environment.yml:356 pins
pytorch 2.5.1, this image has 2.11.0) so the oracle calls the same ATen CPU
ops through ``torch.nn.functional``.  Parameters are a flat dict keyed by the
reference ``state_dict`` names.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import numpy as np
import torch
import torch.nn.functional as F

DEFAULT_NETWORK_CONFIG = {
    # mirrors /root/reference/config/config.json:9-35 ("network" block)
    "encoder": {"in_channels": 2, "num_layers": 2, "kernel_size": 3, "channel_size": 64},
    "recursive": {"alpha_residual": True, "in_channels": 64, "num_layers": 2, "kernel_size": 3},
    "decoder": {
        "deconv": {"in_channels": 64, "kernel_size": 3, "stride": 3, "out_channels": 64},
        "final": {"in_channels": 64, "kernel_size": 1, "out_channels": 1},
    },
}


def param_shapes(config: dict = DEFAULT_NETWORK_CONFIG) -> Dict[str, tuple]:
    """The 31 state_dict tensors of the reference HRNet (HRNet.py:175-184)."""
    enc, rec, dec = config["encoder"], config["recursive"], config["decoder"]
    c, k = enc["channel_size"], enc["kernel_size"]
    shapes = {
        "encode.init_layer.0.weight": (c, enc["in_channels"], k, k),
        "encode.init_layer.0.bias": (c,),
        "encode.init_layer.1.weight": (1,),
    }
    for r in range(enc["num_layers"]):
        for j in (0, 2):
            shapes[f"encode.res_layers.{r}.block.{j}.weight"] = (c, c, k, k)
            shapes[f"encode.res_layers.{r}.block.{j}.bias"] = (c,)
            shapes[f"encode.res_layers.{r}.block.{j + 1}.weight"] = (1,)
    shapes["encode.final.0.weight"] = (c, c, k, k)
    shapes["encode.final.0.bias"] = (c,)
    f, fk = rec["in_channels"], rec["kernel_size"]
    for j in (0, 2):
        shapes[f"fuse.fuse.0.block.{j}.weight"] = (2 * f, 2 * f, fk, fk)
        shapes[f"fuse.fuse.0.block.{j}.bias"] = (2 * f,)
        shapes[f"fuse.fuse.0.block.{j + 1}.weight"] = (1,)
    shapes["fuse.fuse.1.weight"] = (f, 2 * f, fk, fk)
    shapes["fuse.fuse.1.bias"] = (f,)
    shapes["fuse.fuse.2.weight"] = (1,)
    d, fin = dec["deconv"], dec["final"]
    shapes["decode.deconv.0.weight"] = (d["in_channels"], d["out_channels"], d["kernel_size"], d["kernel_size"])
    shapes["decode.deconv.0.bias"] = (d["out_channels"],)
    shapes["decode.deconv.1.weight"] = (1,)
    shapes["decode.final.weight"] = (fin["out_channels"], fin["in_channels"], fin["kernel_size"], fin["kernel_size"])
    shapes["decode.final.bias"] = (fin["out_channels"],)
    return shapes


def make_params(seed: int = 0, config: dict = DEFAULT_NETWORK_CONFIG,
                prelu_jitter: bool = True) -> Dict[str, torch.Tensor]:
    """Deterministic random-init parameters (numpy legacy RandomState, stable
    across machines and torch versions).  Bounds follow torch's default conv
    init (uniform(-1/sqrt(fan_in), 1/sqrt(fan_in))) so activation magnitudes
    match a freshly constructed reference HRNet; PReLU slopes are 0.25 with an
    optional per-layer jitter so a swapped slope is caught by parity tests."""
    rng = np.random.RandomState(seed)
    params = {}
    for name, shape in param_shapes(config).items():
        if len(shape) == 4:
            if name == "decode.deconv.0.weight":      # ConvTranspose2d: (in, out, kh, kw); torch fan_in uses dim 1
                fan_in = shape[1] * shape[2] * shape[3]
            else:
                fan_in = shape[1] * shape[2] * shape[3]
            bound = 1.0 / math.sqrt(fan_in)
            arr = rng.uniform(-bound, bound, size=shape)
            params["_fan_in:" + name.rsplit(".", 1)[0]] = fan_in
        elif shape == (1,):
            arr = np.array([0.25 + (rng.uniform(-0.1, 0.1) if prelu_jitter else 0.0)])
        else:
            fan_in = params.get("_fan_in:" + name.rsplit(".", 1)[0], shape[0])
            bound = 1.0 / math.sqrt(fan_in)
            arr = rng.uniform(-bound, bound, size=shape)
        params[name] = torch.from_numpy(arr.astype(np.float32))
    return {k: v for k, v in params.items() if not k.startswith("_fan_in:")}


def _cast(params, dtype):
    return {k: v.to(dtype) for k, v in params.items()}


def median_anchor(lrs: torch.Tensor) -> torch.Tensor:
    """Lower median over the first min(L, 9) views, zero-padded views included
    (HRNet.py:200).  lrs: (B, L, H, W) -> (B, H, W)."""
    return torch.median(lrs[:, :9], 1).values


def _conv(x, p, key, pad):
    return F.conv2d(x, p[key + ".weight"], p[key + ".bias"], padding=pad)


def _res_block(x, p, prefix, pad):
    """x + PReLU(conv(PReLU(conv(x))))  (HRNet.py:17-33)."""
    y = F.prelu(_conv(x, p, prefix + ".block.0", pad), p[prefix + ".block.1.weight"])
    y = F.prelu(_conv(y, p, prefix + ".block.2", pad), p[prefix + ".block.3.weight"])
    return x + y


def encode(p, x, config=DEFAULT_NETWORK_CONFIG):
    """(N, 2, H, W) -> (N, C, H, W)  (HRNet.py:62-74)."""
    pad = config["encoder"]["kernel_size"] // 2
    x = F.prelu(_conv(x, p, "encode.init_layer.0", pad), p["encode.init_layer.1.weight"])
    for r in range(config["encoder"]["num_layers"]):
        x = _res_block(x, p, f"encode.res_layers.{r}", pad)
    return _conv(x, p, "encode.final.0", pad)


def fuse_level(p, x, alphas, config=DEFAULT_NETWORK_CONFIG):
    """One halving step of HRNet.py:113-132.  x: (B, n, C, H, W), alphas (B, n).
    View i is paired with view n'-1-i (n' = n minus parity; an odd last view is
    dropped)."""
    pad = config["recursive"]["kernel_size"] // 2
    b, n, c, h, w = x.shape
    half = n // 2
    top = n - (n % 2)
    alice = x[:, :half]
    bob = torch.flip(x[:, half:top], [1])
    pair = torch.cat([alice, bob], 2).reshape(b * half, 2 * c, h, w)
    y = _res_block(pair, p, "fuse.fuse.0", pad)
    y = F.prelu(_conv(y, p, "fuse.fuse.1", pad), p["fuse.fuse.2.weight"]).reshape(b, half, c, h, w)
    if config["recursive"]["alpha_residual"]:
        a_bob = torch.flip(alphas[:, half:top], [1]).reshape(b, half, 1, 1, 1)
        y = alice + a_bob * y
        alphas = alphas[:, :half]
    return y, alphas


def fuse(p, x, alphas, config=DEFAULT_NETWORK_CONFIG, trace: Optional[list] = None):
    """(B, L, C, H, W) -> (B, C, H, W)  (HRNet.py:99-134)."""
    while x.shape[1] // 2 > 0:
        x, alphas = fuse_level(p, x, alphas, config)
        if trace is not None:
            trace.append(x)
    return x.mean(1)


def decode(p, x, config=DEFAULT_NETWORK_CONFIG):
    """(B, C, H, W) -> (B, 1, 3H, 3W)  (HRNet.py:158-169)."""
    d = config["decoder"]
    y = F.conv_transpose2d(x, p["decode.deconv.0.weight"], p["decode.deconv.0.bias"], stride=d["deconv"]["stride"])
    y = F.prelu(y, p["decode.deconv.1.weight"])
    return F.conv2d(y, p["decode.final.weight"], p["decode.final.bias"], padding=d["final"]["kernel_size"] // 2)


def hrnet_forward(params, lrs, alphas, config=DEFAULT_NETWORK_CONFIG, dtype=torch.float32,
                  trace: Optional[dict] = None) -> torch.Tensor:
    """Oracle for HRNet.forward(lrs, alphas) (HRNet.py:186-211).  Square inputs
    only: the reference view() at HRNet.py:204 swaps H and W otherwise."""
    p = _cast(params, dtype)
    lrs = torch.as_tensor(lrs).to(dtype)
    alphas = torch.as_tensor(alphas).to(dtype)
    b, l, h, w = lrs.shape
    if h != w:
        raise ValueError("oracle restates the reference for square inputs only (HRNet.py:204)")
    with torch.no_grad():
        anchor = median_anchor(lrs)                                        # (B, H, W)
        stacked = torch.stack([lrs, anchor[:, None].expand(b, l, h, w)], 2)  # (B, L, 2, H, W)
        feats = encode(p, stacked.reshape(b * l, 2, h, w), config)
        feats = feats.reshape(b, l, -1, h, w)
        levels = [] if trace is not None else None
        fused = fuse(p, feats, alphas, config, trace=levels)
        sr = decode(p, fused, config)
    if trace is not None:
        trace.update(anchor=anchor, encoded=feats, levels=levels, fused=fused)
    return sr


def flops_per_imageset(l: int, h: int, w: int) -> float:
    """Algorithmic FLOPs (MAC = 2) per imageset, SURVEY.md section 8d."""
    pairs, n = 0, l
    while n // 2 > 0:
        pairs += n // 2
        n //= 2
    return float(h * w) * (l * 370944.0 + pairs * 737280.0 + 74880.0)
