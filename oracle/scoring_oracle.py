"""Oracle (test infrastructure, not product): numpy restatement of the scoring
half of the hot path.

  * lanczos_taps / lanczos_shift .... /root/reference/src/lanczos.py:5-43, 47-107
    (as called by ShiftNet.transform, ShiftNet.py:87-89: img (1, B*V, H, W),
    shift (B*V, 2) = (dy, dx), a=3, p=5)
  * cpsnr / shift_cpsnr ............. /root/reference/src/Evaluator.py:11-43, 52-73
    (+ get_patch, DataLoader.py:16-30)

Everything is float32 numpy, in the reference's own operation order, so the
cPSNR restatement is bit-identical to the reference on the same numpy build.
``shift_cpsnr`` additionally returns the 49 site scores and the argmax the
reference throws away (Evaluator.py:72 keeps only the max).
"""
from __future__ import annotations

import itertools

import numpy as np


# ----------------------------------------------------------------------------
# Lanczos sub-pixel shift
# ----------------------------------------------------------------------------
def lanczos_taps(d, a: int = 3, n: int = 7) -> np.ndarray:
    """Normalised 1-D Lanczos taps for shift(s) ``d`` -> (..., n) float32.
    w(t) = sinc(pi t) sinc(pi t / a) WITHOUT the |t| < a support clamp and with
    pi*t == 0 replaced by 1e-6 (lanczos.py:26-41)."""
    d = np.asarray(d, dtype=np.float32).reshape(-1, 1)
    lobes = (n - 1) // 2
    x = np.linspace(-lobes, lobes, n, dtype=np.float32).reshape(1, -1) - d
    pix = (np.float32(np.pi) * x).astype(np.float32)
    pix = np.where(pix == 0, np.float32(1e-6), pix).astype(np.float32)
    k = (np.sin(pix) / pix) * (np.sin(pix / np.float32(a)) / (pix / np.float32(a)))
    k = k.astype(np.float32)
    return (k / k.sum(axis=1, keepdims=True)).astype(np.float32)


def _padded_index(q: np.ndarray, size: int, p: int):
    """Index into the original axis for padded-coordinate ``q`` (origin at the
    first real pixel), plus a validity mask: reflect (no edge repeat) inside
    the p-wide ReflectionPad2d ring (lanczos.py:71-72), zero outside it (the
    conv2d zero padding of lanczos.py:90-94)."""
    valid = (q >= -p) & (q < size + p)
    r = np.where(q < 0, -q, q)
    r = np.where(r >= size, 2 * (size - 1) - r, r)
    return np.clip(r, 0, size - 1), valid


def lanczos_shift(img: np.ndarray, shift: np.ndarray, p: int = 3, a: int = 3, n: int = 7) -> np.ndarray:
    """img (Nb, C, H, W), shift (C, 2) = (dy, dx) -> (Nb, C, H, W) float32.
    Closed form of lanczos.py:62-104: reflect-pad by p, correlate with the y
    taps (7x1) then the x taps (1x7), crop p."""
    img = np.asarray(img, dtype=np.float32)
    shift = np.asarray(shift, dtype=np.float32)
    nb, c, h, w = img.shape
    assert shift.shape == (c, 2)
    half = n // 2
    ky = lanczos_taps(shift[:, 0], a, n)   # (C, n)
    kx = lanczos_taps(shift[:, 1], a, n)
    out = np.empty_like(img)
    # columns the x pass reads, in padded coordinates: x + j - half, j in [0, n)
    yy = np.arange(h)
    xx_ext = np.arange(-half, w + half)           # intermediate columns needed by the x pass
    col_idx, col_ok = _padded_index(xx_ext, w, p)
    for ch in range(c):
        tmp = np.zeros((nb, h, w + 2 * half), dtype=np.float32)
        for i in range(n):
            row_idx, row_ok = _padded_index(yy + i - half, h, p)
            rows = img[:, ch][:, row_idx][:, :, col_idx]
            rows = rows * (row_ok[None, :, None] & col_ok[None, None, :])
            tmp += ky[ch, i] * rows
        acc = np.zeros((nb, h, w), dtype=np.float32)
        for j in range(n):
            acc += kx[ch, j] * tmp[:, :, j:j + w]
        out[:, ch] = acc
    return out


def apply_shifts(images: np.ndarray, thetas: np.ndarray) -> np.ndarray:
    """train.py:47-63 + ShiftNet.py:77-90: images (B, V, H, W), thetas (B, V, 2) = (dx, dy); every view becomes one
    channel of a single (1, B*V, H, W) image, the shift is flipped to (dy, dx), a = 3, p = 5."""
    b, v, h, w = images.shape
    flat = np.asarray(images, dtype=np.float32).reshape(1, b * v, h, w)
    shift = np.asarray(thetas, dtype=np.float32).reshape(-1, 2)[:, ::-1]
    return lanczos_shift(flat, shift, p=5, a=3).reshape(b, v, h, w)


# ----------------------------------------------------------------------------
# cPSNR and the 7x7 shift search
# ----------------------------------------------------------------------------
def cpsnr(sr: np.ndarray, hr: np.ndarray, hr_map: np.ndarray):
    """Evaluator.py:11-43 for float inputs in [0, 1] (2-D or (B, n, m))."""
    single = sr.ndim == 2
    if single:
        sr, hr, hr_map = sr[None], hr[None], hr_map[None]
    if sr.dtype.type is np.uint16:
        sr = sr / np.iinfo(np.uint16).max
    else:
        assert 0 <= sr.min() and sr.max() <= 1
    if hr.dtype.type is np.uint16:
        hr = hr / np.iinfo(np.uint16).max
    n_clear = np.sum(hr_map, axis=(1, 2))
    diff = hr - sr
    bias = np.sum(diff * hr_map, axis=(1, 2)) / n_clear
    cmse = np.sum(np.square((diff - bias[:, None, None]) * hr_map), axis=(1, 2)) / n_clear
    out = -10 * np.log10(cmse)
    return out[0] if single else out


def shift_cpsnr(sr: np.ndarray, hr: np.ndarray, hr_map: np.ndarray, border_w: int = 3):
    """Evaluator.py:52-73.  Returns (max, argmax_site, site_scores) where site
    s = x * (2*border_w+1) + y enumerates itertools.product(range(7), range(7))
    (x = row offset of the hr window, y = column offset) and argmax is the first
    maximum like np.argmax.  Batched (B, H, W) inputs give (B,), (B,), (S, B)."""
    size = sr.shape[-1] - 2 * border_w
    src = sr[..., border_w:border_w + size, border_w:border_w + size]
    span = 2 * border_w + 1
    sites = []
    with np.errstate(divide="ignore", invalid="ignore"):
        for x, y in itertools.product(range(span), range(span)):
            sites.append(cpsnr(src, hr[..., x:x + size, y:y + size], hr_map[..., x:x + size, y:y + size]))
    sites = np.array(sites)
    return np.max(sites, axis=0), np.argmax(sites, axis=0), sites


def clear_loss(srs: np.ndarray, hrs: np.ndarray, hr_maps: np.ndarray, metric: str = "cMSE") -> np.ndarray:
    """train.py:66-87 (get_loss) restated in float32 numpy: per-image masked_MSE, cMSE or cPSNR.  Note the weights:
    masked_MSE averages (m*sr - m*hr)^2 over ALL pixels (train.py:78-80); cMSE weights the squared error by m (not m^2,
    unlike Evaluator.cPSNR) and divides by sum(m) (train.py:81-85)."""
    srs, hrs, hr_maps = (np.asarray(x, dtype=np.float32) for x in (srs, hrs, hr_maps))
    if metric == "masked_MSE":
        e = hr_maps * srs - hr_maps * hrs
        return np.mean(e * e, axis=(1, 2), dtype=np.float32)
    nclear = np.sum(hr_maps, axis=(1, 2), dtype=np.float32)
    bright = (np.sum(hr_maps * (hrs - srs), axis=(1, 2), dtype=np.float32) / nclear).astype(np.float32)
    e = (srs + bright[:, None, None]) - hrs
    cmse = np.sum(hr_maps * (e * e), axis=(1, 2), dtype=np.float32) / nclear
    if metric == "cMSE":
        return cmse.astype(np.float32)
    return (-10.0 * np.log10(cmse)).astype(np.float32)


def crop_mask(patch_size: int, crop_size: int) -> np.ndarray:
    """train.py:90-106 (get_crop_mask)."""
    n = 3 * patch_size
    m = np.ones((1, 1, n, n), dtype=np.float32)
    m[0, 0, :crop_size, :] = 0
    m[0, 0, -crop_size:, :] = 0
    m[0, 0, :, :crop_size] = 0
    m[0, 0, :, -crop_size:] = 0
    return m
