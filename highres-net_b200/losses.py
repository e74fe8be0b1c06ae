"""Device twins of the loss helpers of the reference training script (src/train.py:66-106), without autograd:
``get_loss(srs, hrs, hr_maps, metric)`` (masked_MSE / cMSE / cPSNR per image) and ``get_crop_mask``.  (SURVEY.md
section 8f, row N3; the ShiftNet registration network itself stays out of scope.)"""
from __future__ import annotations

import torch

from . import _lib

_METRICS = {"masked_MSE": 0, "cMSE": 1, "cPSNR": 2}


def get_loss(srs, hrs, hr_maps, metric="cMSE"):
    """train.py:66-87.  srs, hrs, hr_maps: CUDA tensors (B, W, H) -> tensor (B,) float32 (no gradient).  Any metric
    other than 'masked_MSE' / 'cMSE' yields cPSNR, like the reference's final return."""
    _lib.require_cuda_tensor(srs, "srs")
    b, h, w = srs.shape
    code = _METRICS.get(metric, 2)
    s = srs.detach().to(torch.float32).contiguous()
    t = hrs.detach().to(device=s.device, dtype=torch.float32).contiguous()
    m = hr_maps.detach().to(device=s.device, dtype=torch.float32).contiguous()
    out = torch.empty(b, dtype=torch.float32, device=s.device)
    with torch.cuda.device(s.device):
        _lib.check(_lib.load().hrn_clear_loss(s.data_ptr(), t.data_ptr(), m.data_ptr(), b, h, w, code, out.data_ptr(),
                                              _lib.current_stream_ptr(s.device)), "hrn_clear_loss")
    return out


def get_crop_mask(patch_size, crop_size):
    """train.py:90-106: (1, 1, 3*patch_size, 3*patch_size) float mask that zeroes a crop_size-wide border."""
    n = 3 * patch_size
    mask = torch.ones((1, 1, n, n), dtype=torch.float32)
    mask[0, 0, :crop_size, :] = 0
    mask[0, 0, -crop_size:, :] = 0
    mask[0, 0, :, :crop_size] = 0
    mask[0, 0, :, -crop_size:] = 0
    return mask
