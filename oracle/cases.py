"""Seeded parity cases shared by oracle/make_golden.py and tests/ (test
infrastructure, not product).  Inputs come from numpy's legacy RandomState so
they are identical in the build container and on the GPU box; only the
reference OUTPUTS are stored under tests/golden/.

Input contract: lrs float32 in [0, 1), padded
trailing views are all-zero with alpha = 0 (utils.py:89-95)."""
from __future__ import annotations

import numpy as np

WEIGHT_SEED = 0

# name -> (B, L, S, number of real views per imageset)
HRNET_CASES = {
    "b2_l4_s32": (2, 4, 32, [4, 2]),        # padded views -> alpha skip path
    "b1_l5_s24": (1, 5, 24, [5]),           # odd L: last view dropped (HRNet.py:115)
    "b1_l6_s16": (1, 6, 16, [6]),           # 6 -> 3 -> 1
    "b2_l9_s16": (2, 9, 16, [9, 7]),        # median over exactly 9 views, 9 -> 4 -> 2 -> 1
    "b1_l16_s16": (1, 16, 16, [13]),        # median over first 9 of 16, four fusion levels
    "b1_l1_s16": (1, 1, 16, [1]),           # no fusion step at all
    "b1_l2_s136": (1, 2, 136, [2]),         # width > 128: two column tiles, ragged second tile
    "c1_b2_l4_s128": (2, 4, 128, [4, 4]),   # BASELINE.json configs[0]
}


def hrnet_inputs(name: str):
    b, l, s, real = HRNET_CASES[name]
    rng = np.random.RandomState(1000 + sorted(HRNET_CASES).index(name))
    lrs = rng.rand(b, l, s, s).astype(np.float32)
    alphas = np.ones((b, l), dtype=np.float32)
    for i, n in enumerate(real):
        lrs[i, n:] = 0.0
        alphas[i, n:] = 0.0
    return lrs, alphas


# name -> (Nb, C, H, W, p, shifts (dy, dx) per channel or None for seeded U(-1, 1))
LANCZOS_CASES = {
    "kat_20x28_p5": (1, 6, 20, 28, 5, [(0.0, 0.0), (1.0, 0.0), (0.0, -1.0), (0.3, -0.7), (-1.0, 1.0), (0.5, 0.5)]),
    "kat_20x28_p3": (1, 6, 20, 28, 3, [(0.0, 0.0), (1.0, 0.0), (0.0, -1.0), (0.3, -0.7), (-1.0, 1.0), (0.5, 0.5)]),
    "batch_2x3_16": (2, 3, 16, 16, 5, None),
    "sr_2_384": (1, 2, 384, 384, 5, None),
    "ragged_3_37x150": (1, 3, 37, 150, 5, None),
}
LANCZOS_TAP_SHIFTS = [0.0, 0.3, 1.0, -0.5, -1.0, 0.999, 2.5]


def lanczos_inputs(name: str):
    nb, c, h, w, p, shifts = LANCZOS_CASES[name]
    rng = np.random.RandomState(2000 + sorted(LANCZOS_CASES).index(name))
    img = rng.rand(nb, c, h, w).astype(np.float32)
    if shifts is None:
        shift = rng.uniform(-1, 1, size=(c, 2)).astype(np.float32)
    else:
        shift = np.array(shifts, dtype=np.float32)
    return img, shift, p


# name -> (B, S, kind)
CPSNR_CASES = {
    "uncorrelated_3_64": (3, 64, "uncorrelated"),   # adversarial: site scores within ~1e-3 dB of each other
    "shifted_5_96": (5, 96, "shifted"),             # hr = roll(sr) + bias + noise: known best site
    "shifted_2_384": (2, 384, "shifted"),           # full Proba-V HR size
    "softmask_2_48": (2, 48, "softmask"),           # non-binary mask (weights m and m^2 differ)
    "degenerate_3_40": (3, 40, "degenerate"),       # all-zero mask -> NaN ; identical -> +inf ; single clear pixel
}


def cpsnr_inputs(name: str):
    b, s, kind = CPSNR_CASES[name]
    rng = np.random.RandomState(3000 + sorted(CPSNR_CASES).index(name))
    sr = rng.rand(b, s, s).astype(np.float32)
    hm = (rng.rand(b, s, s) > 0.1).astype(np.float32)
    if kind == "uncorrelated":
        hr = rng.rand(b, s, s).astype(np.float32)
    elif kind in ("shifted", "softmask"):
        # smooth-ish sr so that a wrong shift is clearly worse, like a real SR image
        sr = (0.1 + 0.05 * sr + 0.3 * np.sin(np.arange(s)[None, :, None] * 0.21 + np.arange(b)[:, None, None])
              * np.cos(np.arange(s)[None, None, :] * 0.13)).astype(np.float32)
        sr = np.clip(sr, 0, 1).astype(np.float32)
        hr = np.empty_like(sr)
        for i in range(b):
            ry, rx = rng.randint(-3, 4, size=2)
            hr[i] = np.roll(sr[i], (ry, rx), axis=(0, 1))
        hr = np.clip(hr + 0.02 + 0.01 * rng.randn(b, s, s), 0, 1).astype(np.float32)
        if kind == "softmask":
            hm = rng.rand(b, s, s).astype(np.float32)
    elif kind == "degenerate":
        hr = rng.rand(b, s, s).astype(np.float32)
        hm[0] = 0.0                      # n_clear = 0 -> NaN everywhere
        hr[1] = sr[1]                    # identical at the centre site -> cMSE = 0 -> +inf
        hm[2] = 0.0
        hm[2, s // 2, s // 2] = 1.0      # one clear pixel: bias = diff, cMSE = 0 at every site containing it
    else:
        raise KeyError(kind)
    return sr, hr, hm


# name -> (number of real views, LR size, has ground truth); min_L = 16 (predict.py:17)
PREDICT_CASES = {
    "v5_s32": (5, 32, True),          # fewer views than min_L: zero padding, alpha = 0
    "v20_s32": (20, 32, True),        # more views than min_L: truncated
    "v16_s24_nohr": (16, 24, False),  # test-split imageset: no HR, score is None
}
PREDICT_MIN_L = 16


def predict_lrs(name: str):
    n, s, _ = PREDICT_CASES[name]
    rng = np.random.RandomState(4000 + sorted(PREDICT_CASES).index(name))
    return rng.rand(n, s, s).astype(np.float32)


# train.get_loss cases: name -> (B, S, kind of mask)
LOSS_CASES = {
    "binary_4_48": (4, 48, "binary"),
    "soft_3_33": (3, 33, "soft"),              # non-binary weights, odd size
    "cropped_2_96": (2, 96, "cropped"),        # hr_map * get_crop_mask(32, 3), as train.py:178-181 uses it
    "full_2_384": (2, 384, "binary"),          # Proba-V HR size
}
LOSS_METRICS = ("masked_MSE", "cMSE", "cPSNR")


def loss_inputs(name: str):
    b, s, kind = LOSS_CASES[name]
    rng = np.random.RandomState(5000 + sorted(LOSS_CASES).index(name))
    sr = rng.rand(b, s, s).astype(np.float32)
    hr = np.clip(sr + 0.03 + 0.05 * rng.randn(b, s, s), 0, 1).astype(np.float32)
    if kind == "soft":
        hm = rng.rand(b, s, s).astype(np.float32)
    else:
        hm = (rng.rand(b, s, s) > 0.15).astype(np.float32)
    if kind == "cropped":
        hm[:, :3, :] = 0
        hm[:, -3:, :] = 0
        hm[:, :, :3] = 0
        hm[:, :, -3:] = 0
    return sr, hr, hm


# train.apply_shifts / ShiftNet.transform cases: name -> (B, V, S)
APPLY_SHIFTS_CASES = {"b2_v3_24": (2, 3, 24), "b1_v1_384": (1, 1, 384), "b3_v16_40": (3, 16, 40)}


def apply_shifts_inputs(name: str):
    b, v, s = APPLY_SHIFTS_CASES[name]
    rng = np.random.RandomState(6000 + sorted(APPLY_SHIFTS_CASES).index(name))
    images = rng.rand(b, v, s, s).astype(np.float32)
    thetas = rng.uniform(-1.5, 1.5, size=(b, v, 2)).astype(np.float32)
    return images, thetas


# predict.evaluate / benchmark: three tiny datasets built from the PREDICT cases.
# name -> (predict case that supplies lr / hr / hr_map, split)
EVALUATE_SETS = {
    "imgset_a": ("v5_s32", "train"),
    "imgset_b": ("v20_s32", "train"),
    "imgset_c": ("v20_s32", "val"),
    "imgset_d": ("v5_s32", "val"),
    "imgset_e": ("v16_s24_nohr", "test"),
}
EVALUATE_BASELINE = {"imgset_a": 40.0, "imgset_b": 45.5, "imgset_c": 50.25, "imgset_d": 38.125, "imgset_e": 47.0}


def evaluate_clearances(name: str):
    n = PREDICT_CASES[EVALUATE_SETS[name][0]][0]
    rng = np.random.RandomState(6000 + sorted(EVALUATE_SETS).index(name))
    return rng.rand(n).astype(np.float64)


def evaluate_lrs(name: str):
    """Distinct inputs per imageset: the predict case's views, scaled a little so that two sets sharing a case differ."""
    scale = 1.0 - 0.05 * sorted(EVALUATE_SETS).index(name)
    return (predict_lrs(EVALUATE_SETS[name][0]) * scale).astype(np.float32)
