"""Generate tests/golden/trainstep.npz: the FORWARD value of the reference's training step (src/train.py:174-186) --
HRNet -> register_batch on the 128 x 128 centre crops -> apply_shifts (Lanczos) -> -get_loss('cPSNR') with the crop
mask -- computed by the unmodified reference functions in eval mode without autograd, on seeded inputs.

Build container only (needs /root/reference):    python oracle/make_golden_trainstep.py"""
from __future__ import annotations

import json
import os
import sys
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import cases, hrnet_oracle, scoring_oracle, shiftnet_oracle  # noqa: E402
from oracle.make_golden import _import_reference  # noqa: E402

B, L, PATCH = 2, 4, 64


def inputs():
    rng = np.random.RandomState(8100)
    yy, xx = np.mgrid[0:PATCH, 0:PATCH].astype(np.float32)
    lrs = np.empty((B, L, PATCH, PATCH), dtype=np.float32)
    for b in range(B):
        base = 0.45 + 0.25 * np.sin(0.21 * xx + b) * np.cos(0.17 * yy - 0.5 * b)
        for v in range(L):
            lrs[b, v] = np.clip(base + 0.03 * rng.randn(PATCH, PATCH), 0, 1)
    return lrs, np.ones((B, L), dtype=np.float32)


def main():
    torch.set_num_threads(os.cpu_count())
    HRNet, _, _ = _import_reference()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from DeepNetworks.ShiftNet import ShiftNet  # type: ignore
        import train as ref_train  # type: ignore
    with open("/root/reference/config/config.json") as f:
        cfg = json.load(f)
    fusion = HRNet(cfg["network"]).eval()
    fusion.load_state_dict(hrnet_oracle.make_params(cases.WEIGHT_SEED), strict=True)
    regis = ShiftNet().eval()
    regis.load_state_dict(shiftnet_oracle.make_params(0), strict=True)
    lrs, alphas = inputs()
    rng = np.random.RandomState(8101)
    crop = 3                                                        # config "crop"
    offset = (3 * PATCH - 128) // 2                                 # train.py:147
    torch_mask = ref_train.get_crop_mask(patch_size=PATCH, crop_size=crop)
    with torch.no_grad():
        srs = fusion(torch.from_numpy(lrs), torch.from_numpy(alphas))                                  # train.py:174
        hr = np.clip(np.roll(srs[:, 0].numpy(), (1, -1), (1, 2)) + 0.02 + 0.01 * rng.randn(B, 3 * PATCH, 3 * PATCH), 0, 1).astype(np.float32)
        hm = (rng.rand(B, 3 * PATCH, 3 * PATCH) > 0.1).astype(np.float32)
        hrs, hr_maps = torch.from_numpy(hr), torch.from_numpy(hm)
        shifts = ref_train.register_batch(regis, srs[:, :, offset:(offset + 128), offset:(offset + 128)],
                                          reference=hrs[:, offset:(offset + 128), offset:(offset + 128)].view(-1, 1, 128, 128))   # :177-179
        srs_shifted = ref_train.apply_shifts(regis, srs, shifts, "cpu")[:, 0]                          # :180
        cropped_mask = torch_mask[0] * hr_maps                                                         # :183
        loss = -ref_train.get_loss(srs_shifted, hrs, cropped_mask, metric="cPSNR")                     # :185
        total = torch.mean(loss) + cfg["training"]["lambda"] * torch.mean(shifts) ** 2                 # :186-187
    # the oracle pieces compose to the same numbers
    o_srs = hrnet_oracle.hrnet_forward(hrnet_oracle.make_params(cases.WEIGHT_SEED), lrs, alphas).numpy()
    o_shifts = shiftnet_oracle.register_batch(shiftnet_oracle.make_params(0), o_srs[:, :, offset:offset + 128, offset:offset + 128],
                                              hr[:, offset:offset + 128, offset:offset + 128][:, None]).numpy()
    assert np.abs(o_srs - srs.numpy()).max() <= 2e-6 and np.abs(o_shifts - shifts.numpy()).max() <= 1e-5
    o_shifted = scoring_oracle.apply_shifts(o_srs, o_shifts)[:, 0]
    o_loss = -scoring_oracle.clear_loss(o_shifted, hr, torch_mask[0].numpy() * hm, "cPSNR")
    assert np.abs(o_shifted - srs_shifted.numpy()).max() <= 1e-5 and np.abs(o_loss - loss.numpy()).max() <= 1e-3
    print("shifts", shifts.numpy().reshape(-1), "loss", loss.numpy(), "total", float(total))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "trainstep.npz"), hr=hr, hr_map=hm, srs=srs.numpy(),
                        shifts=shifts.numpy(), srs_shifted=srs_shifted.numpy(), loss=loss.numpy(), total=np.float32(total),
                        lam=np.float32(cfg["training"]["lambda"]), offset=np.int32(offset), crop=np.int32(crop))


if __name__ == "__main__":
    main()
