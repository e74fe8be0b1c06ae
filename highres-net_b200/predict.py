"""Device-side mirror of the reference's inference caller (the "next" rows N1/N2 of SURVEY.md section 8f):

  * ``collateFunction(min_L)``       -- src/utils.py:49-113: pad / truncate every imageset to ``min_L`` views
                                        (zero views, alpha = 0) and stack the batch; ``collate_device`` does it on the
                                        GPU from the packed real views (``hrn_collate``): no padded bytes over PCIe;
  * ``get_sr_and_score(imset, model, min_L=16)`` -- src/predict.py:17-49: collate, HRNet forward, clip, shifted cPSNR;
  * ``load_model``, ``evaluate``, ``benchmark``, ``Model`` -- src/predict.py:83-158, 200-217: the harness around it
                                        (evaluate batches imagesets of equal shape; file readers / writers stay out).

Same names, arguments and return values as the reference; the forward and the scoring run through the C ABI on
the GPU and the SR never takes the D2H -> NumPy -> cPSNR detour of predict.py:40-45 (it is copied back once, only
because the reference returns it).  ``get_sr_and_score_batch`` scores many imagesets in one forward."""
from __future__ import annotations

from collections.abc import Mapping

import numpy as np
import torch

from . import _lib
from .evaluator import shift_cPSNR


def img_as_float_u16(raw):
    """DataLoader.py:195-198 on the device: ``skimage.img_as_float(lr).astype(np.float32)`` for 16-bit views, i.e.
    x / 65535 with the same rounding.  raw: CUDA uint16 tensor (any shape) -> float32 tensor of that shape."""
    _lib.require_cuda_tensor(raw, "raw")
    if raw.dtype != torch.uint16:
        raise TypeError("img_as_float_u16 expects a uint16 tensor (the on-disk format of the Proba-V views)")
    src = raw.contiguous()
    out = torch.empty(src.shape, dtype=torch.float32, device=src.device)
    with torch.cuda.device(src.device):
        _lib.check(_lib.load().hrn_u16_to_unit_float(src.data_ptr(), src.numel(), out.data_ptr(),
                                                     _lib.current_stream_ptr(src.device)), "hrn_u16_to_unit_float")
    return out


def img_as_uint_u16(sr):
    """predict.py:176 on the device: ``skimage.img_as_uint(sr)`` for a float32 image, i.e. clip(rint(x * 65535), 0, 65535)
    in fp32 with round-half-to-even (scikit-image 0.24 ``util/dtype.py: _convert``).  sr: CUDA float32 tensor (any shape)
    -> uint16 tensor of that shape (half the D2H bytes of the float image).  Raises ValueError like skimage when a value
    lies outside [-1, 1] (an unclipped SR image can: clip first, as predict.py:43 does for scoring)."""
    _lib.require_cuda_tensor(sr, "sr")
    if sr.dtype != torch.float32:
        raise TypeError("img_as_uint_u16 expects a float32 tensor")
    src = sr.contiguous()
    out = torch.empty(src.shape, dtype=torch.uint16, device=src.device)
    bad = torch.zeros(1, dtype=torch.int32, device=src.device)
    with torch.cuda.device(src.device):
        _lib.check(_lib.load().hrn_unit_float_to_u16(src.data_ptr(), src.numel(), out.data_ptr(), bad.data_ptr(),
                                                     _lib.current_stream_ptr(src.device)), "hrn_unit_float_to_u16")
    if int(bad.item()) != 0:
        raise ValueError("Images of type float must be between -1 and 1.")
    return out


def _packed_views(batch, min_L):
    """The real views of a batch, truncated to min_L per imageset and packed back to back: (packed (sum_b n_b, H, W) host
    tensor in the views' own dtype -- float32 or raw uint16 --, offsets (B + 1) int32)."""
    views = [torch.as_tensor(imageset["lr"])[:min_L] for imageset in batch]
    shapes = {tuple(v.shape[1:]) for v in views}
    dtypes = {v.dtype for v in views}
    if len(shapes) != 1 or len(dtypes) != 1:
        raise ValueError("collate: all imagesets of a batch must share one view size and dtype")
    dtype = dtypes.pop()
    if dtype not in (torch.float32, torch.uint16):
        views, dtype = [v.to(torch.float32) for v in views], torch.float32
    counts = torch.tensor([0] + [v.shape[0] for v in views], dtype=torch.int32)
    packed = torch.empty((int(counts.sum()),) + shapes.pop(), dtype=dtype, pin_memory=torch.cuda.is_available())
    torch.cat(views, dim=0, out=packed)
    return packed, torch.cumsum(counts, 0, dtype=torch.int32)


def _targets(batch):
    """hr / hr_map / names of a batch with the reference's rule (utils.py:97-111): as soon as one imageset has no 'hr',
    the hr list stops growing and neither list is stacked."""
    names = [imageset["name"] for imageset in batch]
    maps = [torch.as_tensor(imageset["hr_map"]) if imageset["hr_map"] is not None else None for imageset in batch]
    hrs = []
    for imageset in batch:
        if imageset["hr"] is None:
            return hrs, maps, names
        hrs.append(torch.as_tensor(imageset["hr"]))
    return torch.stack(hrs, dim=0), torch.stack(maps, dim=0), names


def collate_device(batch, min_L, device):
    """utils.collateFunction (utils.py:63-113) ON THE DEVICE: only the real views are copied to the GPU (one packed,
    pinned staging buffer, one H2D copy; float32, or raw uint16 at half the bytes), and ``hrn_collate`` scatters them
    into the padded (B, min_L, H, W) float32 batch, zero-fills the padding planes and writes the alphas.  Returns
    (lrs, alphas) on ``device`` plus (hr_batch, hm_batch, names) exactly like ``collateFunction``."""
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("collate_device needs a CUDA device: the B200 path has no CPU fallback")
    packed, offsets = _packed_views(batch, min_L)
    b, (h, w) = len(batch), packed.shape[1:]
    with torch.cuda.device(device):
        packed_dev = packed.to(device, non_blocking=True)
        offsets_dev = offsets.to(device, non_blocking=True)
        lrs = torch.empty((b, min_L, h, w), dtype=torch.float32, device=device)
        alphas = torch.empty((b, min_L), dtype=torch.float32, device=device)
        _lib.check(_lib.load().hrn_collate(packed_dev.data_ptr(), int(packed.dtype == torch.uint16), offsets_dev.data_ptr(),
                                           b, int(min_L), h, w, lrs.data_ptr(), alphas.data_ptr(),
                                           _lib.current_stream_ptr(device)), "hrn_collate")
        packed_dev.record_stream(torch.cuda.current_stream(device))
    hrs, hr_maps, names = _targets(batch)
    return lrs, alphas, hrs, hr_maps, names


class collateFunction:
    """Util class to create padded batches of data (utils.py:49-113): same constructor, call convention and return
    values.  With ``device=None`` (the reference's signature) the padded batch is built on the host -- one zero-initialised
    (B, min_L, H, W) tensor that the real views are written into; with a CUDA ``device`` the batch is built by
    ``collate_device`` and never exists on the host."""

    def __init__(self, min_L=32, device=None):
        self.min_L = min_L
        self.device = device

    def __call__(self, batch):
        return self.collateFunction(batch)

    def collateFunction(self, batch):
        """batch: list of imagesets (mappings with 'lr' (L, H, W), 'hr', 'hr_map', 'name') ->
        (padded_lr_batch (B, min_L, H, W), alpha_batch (B, min_L), hr_batch, hm_batch, names)."""
        if self.device is not None:
            return collate_device(batch, self.min_L, self.device)
        packed, offsets = _packed_views(batch, self.min_L)
        if packed.dtype == torch.uint16:
            packed = packed.to(torch.float32) / 65535.0                 # DataLoader.py:195-198
        offsets = offsets.tolist()
        lrs = torch.zeros((len(batch), self.min_L) + tuple(packed.shape[1:]), dtype=packed.dtype)
        alphas = torch.zeros((len(batch), self.min_L))
        for i in range(len(batch)):
            n = offsets[i + 1] - offsets[i]
            lrs[i, :n] = packed[offsets[i]:offsets[i + 1]]
            alphas[i, :n] = 1.0
        hrs, hr_maps, names = _targets(batch)
        return lrs, alphas, hrs, hr_maps, names


def _device_of(model):
    return next(model.parameters()).device


def get_sr_and_score(imset, model, min_L=16):
    """predict.py:17-49.  imset: one imageset (mapping) or a tuple of batches (lrs, alphas, hrs, hr_maps, names).
    Returns (sr: np.ndarray (3H, 3W) of the FIRST imageset, scPSNR: float or None)."""
    device = _device_of(model)
    if device.type != "cuda":
        raise RuntimeError("the model must live on a CUDA device: the B200 path has no CPU fallback")
    if isinstance(imset, Mapping):
        lrs, alphas, hrs, hr_maps, names = collate_device([imset], min_L, device)
    elif isinstance(imset, tuple):
        lrs, alphas, hrs, hr_maps, names = imset
    else:
        raise TypeError("imset must be an imageset mapping or a tuple of batches")
    sr_dev = model(lrs.float().to(device), alphas.float().to(device))[:, 0]
    sr = sr_dev.detach().cpu().numpy()[0]
    if len(hrs) > 0:
        hr = torch.as_tensor(hrs)[0:1].float().to(device)
        hm = torch.as_tensor(hr_maps)[0:1].float().to(device)
        score = shift_cPSNR(sr_dev[0:1], hr, hm, border_w=3, clip_sr=True)       # np.clip(sr, 0, 1) fused (predict.py:43)
        sc_psnr = np.float32(score[0].item())
    else:
        sc_psnr = None
    return sr, sc_psnr


def get_sr_and_score_batch(imsets, model, min_L=16):
    """New capability: all imagesets in ONE forward + ONE scoring launch.
    Returns (srs: np.ndarray (B, 3H, 3W), scores: np.ndarray (B,) float32 or None when there is no ground truth)."""
    device = _device_of(model)
    if device.type != "cuda":
        raise RuntimeError("the model must live on a CUDA device: the B200 path has no CPU fallback")
    lrs, alphas, hrs, hr_maps, names = collate_device(list(imsets), min_L, device)
    sr_dev = model(lrs, alphas)[:, 0]
    scores = None
    if len(hrs) > 0:
        scores = shift_cPSNR(sr_dev, hrs.float().to(device), hr_maps.float().to(device), border_w=3, clip_sr=True)
        scores = scores.cpu().numpy().astype(np.float32)
    return sr_dev.cpu().numpy(), scores


# ------------------------------------------------------------------ the harness around get_sr_and_score
def load_model(config, checkpoint_file, device=None):
    """predict.py:83-100: build ``HRNet(config["network"])`` on the GPU and load a checkpoint written by the reference
    (``torch.save(fusion_model.state_dict(), ...)``, train.py:220-222 -- the 31 state_dict keys are the same)."""
    from .hrnet import HRNet
    if device is None:
        if not torch.cuda.is_available():
            raise RuntimeError("load_model: no CUDA device; the B200 path has no CPU fallback")
        device = torch.device("cuda", torch.cuda.current_device())
    model = HRNet(config["network"]).to(device)
    model.load_state_dict(torch.load(checkpoint_file, map_location="cpu"))
    # The reference hands the module back in training mode and calls it that way (predict.py:39; HRNet has no dropout or
    # batch norm, so the mode changes nothing there).  This path is inference-only and refuses training mode with grad
    # enabled, so the loaded model is switched to eval here.
    return model.eval()


def _shape_key(imset):
    return (tuple(imset["lr"].shape[1:]), imset["hr"] is not None)


def evaluate(model, train_dataset, val_dataset, test_dataset, min_L=16, batch_size=32):
    """predict.py:103-135: scores every imageset of the three datasets.  Returns the same three dicts
    (scores, clerances, part), keyed by imageset name.  Unlike the reference loop (one imageset per forward, D2H of the
    SR image, NumPy scoring) imagesets of equal shape are grouped ``batch_size`` at a time into one forward and one
    scoring launch; the SR image never leaves the device.  Per-imageset results are identical to get_sr_and_score
    (imagesets are independent in every kernel)."""
    model.eval()
    device = _device_of(model)
    if device.type != "cuda":
        raise RuntimeError("the model must live on a CUDA device: the B200 path has no CPU fallback")
    scores, clerances, part = {}, {}, {}
    for split, dataset in (("train", train_dataset), ("val", val_dataset), ("test", test_dataset)):
        pending = {}                                   # shape key -> imagesets waiting for a full batch

        def flush(group):
            lrs, alphas, hrs, hr_maps, names = collate_device(group, min_L, device)
            sr_dev = model(lrs, alphas)[:, 0]
            if len(hrs) > 0:
                sc = shift_cPSNR(sr_dev, hrs.float().to(device), hr_maps.float().to(device), border_w=3, clip_sr=True)
                sc = sc.cpu().numpy().astype(np.float32)
            for i, imset in enumerate(group):
                scores[imset["name"]] = sc[i] if len(hrs) > 0 else None
                clerances[imset["name"]] = imset["clearances"] if "clearances" in imset else None
                part[imset["name"]] = split

        for imset in dataset if dataset is not None else ():
            key = _shape_key(imset)
            group = pending.setdefault(key, [])
            group.append(imset)
            if len(group) == batch_size:
                flush(group)
                pending[key] = []
        for group in pending.values():
            if group:
                flush(group)
    return scores, clerances, part


def benchmark(baseline_cpsnrs, scores, part, clerances):
    """predict.py:138-158: table of the ESA baseline against the model (score = ESA / model, clearance statistics)."""
    import pandas as pd
    table = pd.DataFrame({"ESA": baseline_cpsnrs, "model": scores, "clr": clerances, "part": part})
    table["score"] = table["ESA"] / table["model"]
    table["mean_clr"] = table["clr"].map(np.mean)
    table["std_clr"] = table["clr"].map(np.std)
    return table


class Model(object):
    """predict.py:200-220 without the file writers: ``Model(config).load_checkpoint(f)``, ``model(imset)`` ->
    (sr, scPSNR), ``model.evaluate(train, val, test, baseline_cpsnrs)`` -> results table."""

    def __init__(self, config):
        self.config = config
        self.model = None

    def load_checkpoint(self, checkpoint_file):
        self.model = load_model(self.config, checkpoint_file)

    def __call__(self, imset):
        return get_sr_and_score(imset, self.model, min_L=self.config["training"]["min_L"])

    def evaluate(self, train_dataset, val_dataset, test_dataset, baseline_cpsnrs):
        scores, clearance, part = evaluate(self.model, train_dataset, val_dataset, test_dataset,
                                           min_L=self.config["training"]["min_L"])
        return benchmark(baseline_cpsnrs, scores, part, clearance)
