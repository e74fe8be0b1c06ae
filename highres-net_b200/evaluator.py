"""Drop-in for the reference ``Evaluator`` module (src/Evaluator.py): ``cPSNR`` and
``shift_cPSNR`` with the same arguments and return types (numpy in -> numpy
float32 out), computed by the CUDA shift-search kernels.  CUDA tensors are
accepted too (and then returned), which removes the D2H copy of train.py:208.
``shift_cPSNR_argmax`` additionally returns the best shift the reference computes
but discards (Evaluator.py:72)."""
from __future__ import annotations

import numpy as np
import torch

from . import _lib


def _to_device(x, device):
    if torch.is_tensor(x):
        t = x.detach()
        if not t.is_cuda:
            t = t.to(device)
    else:
        arr = np.asarray(x)
        if arr.dtype == np.uint16:                      # Evaluator.py:27-32: uint16 means [0, 65535]
            arr = arr / np.iinfo(np.uint16).max
        t = torch.from_numpy(np.ascontiguousarray(arr, dtype=np.float32)).to(device)
    return t.to(torch.float32).contiguous()


def _search(sr, hr, hr_map, border_w, clip_sr=False, device=None):
    is_numpy = not torch.is_tensor(sr)
    if device is None:
        device = sr.device if (torch.is_tensor(sr) and sr.is_cuda) else torch.device("cuda", torch.cuda.current_device())
    if is_numpy and np.asarray(sr).dtype != np.uint16 and not clip_sr:
        s = np.asarray(sr)
        assert 0 <= s.min() and s.max() <= 1, 'sr.dtype must be either uint16 (range 0-65536) or float64 in (0, 1).'
    srd, hrd, hmd = (_to_device(t, device) for t in (sr, hr, hr_map))
    single = srd.dim() == 2
    if single:
        srd, hrd, hmd = srd[None], hrd[None], hmd[None]
    b, h, w = srd.shape
    sites = (2 * border_w + 1) ** 2
    best = torch.empty(b, dtype=torch.float32, device=device)
    arg = torch.empty(b, dtype=torch.int32, device=device)
    table = torch.empty((b, sites), dtype=torch.float32, device=device)
    with torch.cuda.device(device):
        _lib.check(_lib.load().hrn_shift_cpsnr(srd.data_ptr(), hrd.data_ptr(), hmd.data_ptr(), b, h, w, int(border_w),
                                               int(bool(clip_sr)), best.data_ptr(), arg.data_ptr(), table.data_ptr(),
                                               _lib.current_stream_ptr(device)), "hrn_shift_cpsnr")
    return best, arg, table, single, is_numpy


def scoring_debug_set(knob: str, value: int) -> None:
    """Process-wide test knobs of the scoring kernels (hrn_scoring_debug_set), e.g. ("cpsnr_generic", 1)."""
    _lib.check(_lib.load().hrn_scoring_debug_set(knob.encode(), int(value)), "hrn_scoring_debug_set")


def _finish(t, single, is_numpy):
    if is_numpy:
        a = t.cpu().numpy()
        return a[0] if single else a
    return t[0] if single else t


def cPSNR(sr, hr, hr_map):
    """Evaluator.py:11-43: brightness-bias-corrected clear PSNR of (n, m) or (B, n, m) images."""
    best, _, _, single, is_numpy = _search(sr, hr, hr_map, border_w=0)
    return _finish(best, single, is_numpy)


def shift_cPSNR(sr, hr, hr_map, border_w=3, clip_sr=False):
    """Evaluator.py:52-73: max cPSNR over the (2*border_w+1)^2 integer shifts of hr against the cropped sr."""
    best, _, _, single, is_numpy = _search(sr, hr, hr_map, border_w, clip_sr)
    return _finish(best, single, is_numpy)


def shift_cPSNR_argmax(sr, hr, hr_map, border_w=3, clip_sr=False):
    """Like shift_cPSNR but returns (max_cPSNR, (x, y), site_scores): (x, y) is the winning hr window offset in
    itertools.product(range(2b+1), range(2b+1)) order (first maximum, np.argmax), site_scores has (2b+1)^2 entries."""
    best, arg, table, single, is_numpy = _search(sr, hr, hr_map, border_w, clip_sr)
    span = 2 * border_w + 1
    xy = torch.stack([arg // span, arg % span], dim=-1)
    return _finish(best, single, is_numpy), _finish(xy, single, is_numpy), _finish(table, single, is_numpy)
