"""Oracle (test infrastructure, not product): restatement of the reference inference caller.

  * collate ............ /root/reference/src/utils.py:63-113 (collateFunction)
  * get_sr_and_score ... /root/reference/src/predict.py:17-49

built on the other oracle pieces (hrnet_oracle.hrnet_forward, scoring_oracle.shift_cpsnr)."""
from __future__ import annotations

import numpy as np

from . import hrnet_oracle, scoring_oracle


def collate(imsets, min_l):
    """-> lrs (B, min_l, H, W) float32, alphas (B, min_l) float32; zero-padded / truncated views."""
    lrs, alphas = [], []
    for im in imsets:
        lr = np.asarray(im["lr"], dtype=np.float32)
        n = lr.shape[0]
        if n >= min_l:
            lrs.append(lr[:min_l])
            alphas.append(np.ones(min_l, dtype=np.float32))
        else:
            lrs.append(np.concatenate([lr, np.zeros((min_l - n,) + lr.shape[1:], dtype=np.float32)], 0))
            alphas.append(np.concatenate([np.ones(n, dtype=np.float32), np.zeros(min_l - n, dtype=np.float32)]))
    return np.stack(lrs), np.stack(alphas)


def get_sr_and_score(imset, params, min_l=16):
    lrs, alphas = collate([imset], min_l)
    sr = hrnet_oracle.hrnet_forward(params, lrs, alphas).numpy()[0, 0]
    if imset["hr"] is None:
        return sr, None
    score = scoring_oracle.shift_cpsnr(np.clip(sr, 0, 1), np.asarray(imset["hr"], np.float32),
                                       np.asarray(imset["hr_map"], np.float32))[0]
    return sr, np.float32(score)


def img_as_uint(image):
    """skimage.img_as_uint for a float32 image (predict.py:176), restated from scikit-image 0.24.0 (the version
    environment.yml:375 pins; the package is not installed in the build container, so this one function is NOT pinned
    against the library itself): util/dtype.py `_convert`, float -> unsigned branch: values outside [-1, 1] raise;
    image * 65535 in float32 (the computation type for a 2-byte output), np.rint (half to even), clip to [0, 65535]."""
    image = np.asarray(image, dtype=np.float32)
    if np.min(image) < -1.0 or np.max(image) > 1.0:
        raise ValueError("Images of type float must be between -1 and 1.")
    out = np.multiply(image, 65535, dtype=np.float32)
    np.rint(out, out=out)
    np.clip(out, 0, 65535, out=out)
    return out.astype(np.uint16)
