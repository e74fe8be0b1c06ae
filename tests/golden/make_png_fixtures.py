"""Writes the synthetic Proba-V-shaped imageset fixture under tests/golden/imgsets/ (run once, in the build container):

    python tests/golden/make_png_fixtures.py

Files are written with PIL -- the library behind skimage.io.imread / imsave (imageio -> pillow), which the reference uses
(DataLoader.py:134-140, predict.py:181) -- so the native PNG reader is checked against files the reference's own stack
produces.  Layout per imageset (DataLoader.py:107-140): LRnnn.png (16-bit grey views), QMnnn.png (8-bit status maps, 0/255),
SM.png (8-bit HR status map), HR.png (16-bit, absent in the 'test' set).  imgsets.npz holds the same pixels as arrays."""
import os

import numpy as np
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "imgsets")
SETS = {  # name -> (channel dir, views, has HR)
    "imgset0001": ("RED", 5, True),
    "imgset0002": ("RED", 3, True),
    "imgset0003": ("NIR", 4, False),
}
S = 24


def main():
    rng = np.random.RandomState(2024)
    arrays = {}
    for name, (chan, views, has_hr) in SETS.items():
        d = os.path.join(OUT, chan, name)
        os.makedirs(d, exist_ok=True)
        base = rng.rand(3 * S, 3 * S)
        for v in range(views):
            lr = base[v % 3::3, (v // 3) % 3::3][:S, :S] * 0.6 + 0.05 * rng.rand(S, S)      # smooth content + noise
            lr16 = np.round(lr * 65535).astype(np.uint16)
            qm = ((rng.rand(S, S) > 0.1 * (v + 1) / views) * 255).astype(np.uint8)
            if name == "imgset0001" and v in (1, 3):
                qm[:] = 255                                   # two equally (fully) clear views: the tie case of the sort
            Image.fromarray(lr16).save(os.path.join(d, f"LR{v:03d}.png"), optimize=bool(v % 2))
            Image.fromarray(qm).save(os.path.join(d, f"QM{v:03d}.png"))
            arrays[f"{name}/LR{v:03d}"] = lr16
            arrays[f"{name}/QM{v:03d}"] = qm
        sm = ((rng.rand(3 * S, 3 * S) > 0.1) * 255).astype(np.uint8)
        Image.fromarray(sm).save(os.path.join(d, "SM.png"))
        arrays[f"{name}/SM"] = sm
        if has_hr:
            hr16 = np.round(base * 0.6 * 65535).astype(np.uint16)
            Image.fromarray(hr16).save(os.path.join(d, "HR.png"))
            arrays[f"{name}/HR"] = hr16
    # one 1-bit file as well (PIL mode "1"): status maps are sometimes stored that way
    bit = rng.rand(S, S + 3) > 0.5
    Image.fromarray(bit).save(os.path.join(OUT, "bits.png"))
    arrays["bits"] = bit
    np.savez_compressed(os.path.join(HERE, "imgsets.npz"), **arrays)


if __name__ == "__main__":
    main()
