"""Drop-in for the reference ``lanczos`` module (src/lanczos.py): ``lanczos_kernel``
and ``lanczos_shift`` with the same arguments, run by one bandwidth-bound CUDA
kernel (reflect padding, tap generation and both separable passes fused) instead
of a Python loop of ~25 tiny launches per image.  CUDA tensors only."""
from __future__ import annotations

import torch

from . import _lib


def lanczos_kernel(dx, a=3, N=7, dtype=None, device=None):
    """lanczos.py:5-43.  dx: tensor (n, 1) (or anything torch.tensor accepts) -> (n, N) normalised taps."""
    if not torch.is_tensor(dx):
        dx = torch.tensor(dx, dtype=dtype, device=device)
    if device is None:
        device = dx.device
    if dtype is None:
        dtype = dx.dtype
    d = dx.to(device=device, dtype=torch.float32).contiguous().reshape(-1)
    _lib.require_cuda_tensor(d, "dx")
    out = torch.empty((d.numel(), N), dtype=torch.float32, device=d.device)
    with torch.cuda.device(d.device):
        _lib.check(_lib.load().hrn_lanczos_taps(d.data_ptr(), d.numel(), int(a), int(N), out.data_ptr(),
                                                _lib.current_stream_ptr(d.device)), "hrn_lanczos_taps")
    return out.to(dtype)


def lanczos_shift(img, shift, p=3, a=3, N=7):
    """lanczos.py:47-107.  img (batch, channels, H, W), shift (channels, 2) = (dy, dx) per channel."""
    _lib.require_cuda_tensor(img, "img")
    if img.dim() != 4:
        raise ValueError("img must be (batch_size, channels, height, width)")
    nb, c, h, w = img.shape
    shift = torch.as_tensor(shift, device=img.device).to(torch.float32).contiguous()
    if tuple(shift.shape) != (c, 2):
        raise ValueError("shift must be (channels, 2)")
    src = img.detach().to(torch.float32).contiguous()
    out = torch.empty_like(src)
    with torch.cuda.device(img.device):
        _lib.check(_lib.load().hrn_lanczos_shift(src.data_ptr(), shift.data_ptr(), nb, c, h, w, int(p), int(a), int(N),
                                                 out.data_ptr(), _lib.current_stream_ptr(img.device)),
                   "hrn_lanczos_shift")
    return out.to(img.dtype)


def transform(theta, I, device=None):
    """ShiftNet.transform (ShiftNet.py:77-90): shift image b of I (B, 1, H, W) by theta[b] = (dx, dy) with Lanczos
    interpolation (a = 3, p = 5) -> (1, 1, B, H, W), exactly the reference's (odd) result shape.  ``device`` is accepted
    for signature compatibility and ignored: the work happens on I.device."""
    _lib.require_cuda_tensor(I, "I")
    theta = torch.as_tensor(theta, device=I.device)
    return lanczos_shift(I.transpose(0, 1), theta.flip(-1), p=5, a=3)[:, None]


def apply_shifts(shiftNet, images, thetas, device=None):
    """train.apply_shifts (train.py:47-63): images (B, V, H, W), thetas (B, V, 2) = (dx, dy) -> warped (B, V, H, W).
    ``shiftNet`` is only used by the reference to reach ``transform``; pass None (or any object) here."""
    batch_size, n_views, height, width = images.shape
    flat = images.reshape(-1, 1, height, width)
    new_images = transform(thetas.reshape(-1, 2), flat, device=device)
    return new_images.view(-1, n_views, height, width)
