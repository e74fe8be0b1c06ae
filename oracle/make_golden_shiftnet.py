"""Generate tests/golden/shiftnet.npz by running the UNMODIFIED reference ShiftNet (src/DeepNetworks/ShiftNet.py) in eval
mode and train.register_batch (src/train.py:26-44) on the seeded parameters / pairs of oracle/shiftnet_oracle.py.

Build container only (needs /root/reference):    python oracle/make_golden_shiftnet.py"""
from __future__ import annotations

import os
import sys
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import shiftnet_oracle  # noqa: E402
from oracle.make_golden import _import_reference  # noqa: E402

N_PAIRS, B, V = 6, 2, 3


def main():
    torch.set_num_threads(os.cpu_count())
    _import_reference()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from DeepNetworks.ShiftNet import ShiftNet  # type: ignore
        import train as ref_train  # type: ignore
    params = shiftnet_oracle.make_params(0)
    model = ShiftNet().eval()
    model.load_state_dict(params, strict=True)
    out = {"state_dict_keys": np.array(list(model.state_dict().keys())),
           "n_params": np.int64(sum(p.numel() for p in model.parameters()))}
    x = shiftnet_oracle.make_pairs(N_PAIRS, 0)
    with torch.no_grad():
        ref = model(torch.from_numpy(x)).numpy()
        ref64 = model.double()(torch.from_numpy(x).double()).numpy()
        model.float()
    mine = shiftnet_oracle.shiftnet_forward(params, x).numpy()
    print("theta reference:\n", ref)
    print(f"oracle-vs-reference max|d| = {np.abs(mine - ref).max():.3e}, fp32-vs-fp64 reference = {np.abs(ref - ref64).max():.3e}")
    assert np.abs(mine - ref).max() <= 1e-5 * max(1.0, np.abs(ref).max())
    out["theta"] = ref.astype(np.float32)
    out["theta_fp64"] = ref64.astype(np.float64)
    # register_batch: B imagesets x V views against one reference crop each
    pairs = shiftnet_oracle.make_pairs(B * V, 1).reshape(B, V, 2, 128, 128)
    lrs, reference = pairs[:, :, 1], pairs[:, 0, 0][:, None]
    with torch.no_grad():
        thetas = ref_train.register_batch(model, torch.from_numpy(lrs), torch.from_numpy(reference)).numpy()
    o = shiftnet_oracle.register_batch(params, lrs, reference).numpy()
    assert thetas.shape == (B, V, 2) and np.abs(o - thetas).max() <= 1e-5 * max(1.0, np.abs(thetas).max())
    out["register_thetas"] = thetas.astype(np.float32)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "shiftnet.npz"), **out)
    print("register_batch thetas:\n", thetas)


if __name__ == "__main__":
    main()
