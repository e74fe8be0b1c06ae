"""The file formats on both ends of the path, mirroring the reference's names:

  * ``read_imageset`` / ``ImagesetDataset``      -- src/DataLoader.py:73-148, 153-205: LR*.png / HR.png / SM.png / clearance.npy
                                                   of one imageset directory, views in clearance order (or sampled);
  * ``sample_clearest``, ``get_patch``           -- src/DataLoader.py:16-30, 41-70;
  * ``save_clearance_scores``                    -- src/save_clearance.py:13-27;
  * ``generate_submission_file``                 -- src/predict.py:161-194: SR -> img_as_uint -> 16-bit PNG -> submission.zip.

The PNG codec, the clearance sums / order and the stored ZIP live in the native library (csrc/imageset_io.cu: a thread pool
decodes all views of an imageset at once into ONE pinned uint16 buffer).  With ``raw16=True`` the views stay in their
16-bit on-disk format all the way to the GPU (half the H2D bytes; ``hrn_collate`` / ``hrn_forward_host_u16`` scale them on
the device exactly like ``skimage.img_as_float(...).astype(float32)``, DataLoader.py:195-198); the default returns the same
float32 tensors as the reference.  scikit-image is not needed."""
from __future__ import annotations

import ctypes
import glob
import os
from collections import OrderedDict
from os.path import basename, exists, isfile, join

import numpy as np
import torch

from . import _lib


def _c_paths(paths):
    arr = (ctypes.c_char_p * len(paths))()
    arr[:] = [os.fsencode(p) for p in paths]
    return arr


def png_info(path):
    """-> (height, width, bit_depth, color_type) of a PNG file (hrn_png_info)."""
    w, h, d, c = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32()
    _lib.check(_lib.load().hrn_png_info(os.fsencode(path), ctypes.byref(w), ctypes.byref(h), ctypes.byref(d), ctypes.byref(c)),
               "hrn_png_info")
    return h.value, w.value, d.value, c.value


def read_png_u16(paths, threads=0, pin=None):
    """Decodes greyscale PNG files of one common size on a native thread pool -> uint16 tensor (n, H, W) with the sample
    values as stored.  ``pin`` (default: when CUDA is available) puts the result in pinned memory for async H2D copies."""
    paths = list(paths)
    if not paths:
        raise ValueError("read_png_u16: no files")
    h, w, _, _ = png_info(paths[0])
    pin = torch.cuda.is_available() if pin is None else pin
    out = torch.empty((len(paths), h, w), dtype=torch.uint16, pin_memory=bool(pin))
    _lib.check(_lib.load().hrn_png_read_gray_u16(_c_paths(paths), len(paths), h, w, out.data_ptr(), int(threads)),
               "hrn_png_read_gray_u16")
    return out


def write_png_u16(paths, images, threads=0):
    """images: uint16 array / CPU tensor (n, H, W) (or (H, W) with one path) -> 16-bit greyscale PNG files."""
    paths = [paths] if isinstance(paths, (str, os.PathLike)) else list(paths)
    arr = images.numpy() if torch.is_tensor(images) else np.asarray(images)
    if arr.dtype != np.uint16:
        raise TypeError("write_png_u16 expects uint16 images (use predict.img_as_uint_u16 on the SR image)")
    if arr.ndim == 2:
        arr = arr[None]
    arr = np.ascontiguousarray(arr)
    if arr.shape[0] != len(paths):
        raise ValueError("one path per image")
    _lib.check(_lib.load().hrn_png_write_gray_u16(_c_paths(paths), len(paths), arr.shape[1], arr.shape[2],
                                                  arr.ctypes.data_as(ctypes.c_void_p), int(threads)),
               "hrn_png_write_gray_u16")


def clearance_order(clearances):
    """DataLoader.py:128: ``np.argsort(clearances)[::-1]`` (equally clear views by descending index)."""
    c = np.ascontiguousarray(clearances, dtype=np.float64)
    order = np.empty(len(c), dtype=np.int32)
    _lib.check(_lib.load().hrn_clearance_order(c.ctypes.data_as(ctypes.c_void_p), len(c), order.ctypes.data_as(ctypes.c_void_p)),
               "hrn_clearance_order")
    return order.astype(np.int64)


def _view_ids(imset_dir):
    return np.sort(np.array([basename(path)[2:-4] for path in glob.glob(join(imset_dir, "QM*.png"))]))


def save_clearance_scores(dataset_directories, threads=0):
    """save_clearance.py:13-27: per imageset directory, the sum of every QM status map -> ``clearance.npy``."""
    for imset_dir in dataset_directories:
        ids = _view_ids(imset_dir)
        paths = [join(imset_dir, f"QM{i}.png") for i in ids]
        h, w, _, _ = png_info(paths[0])
        scores = np.empty(len(paths), dtype=np.float64)
        _lib.check(_lib.load().hrn_clearance_scores(_c_paths(paths), len(paths), h, w, int(threads),
                                                    scores.ctypes.data_as(ctypes.c_void_p)), "hrn_clearance_scores")
        np.save(join(imset_dir, "clearance.npy"), scores.astype(np.uint64))      # lr_maps.sum(axis=(1, 2)) of uint16 maps


def get_patch(img, x, y, size=32):
    """DataLoader.py:16-30: square patch with top-left corner (x, y), broadcast over leading dimensions."""
    return img[..., x:(x + size), y:(y + size)]


def sample_clearest(clearances, n=None, beta=50, seed=None):
    """DataLoader.py:41-70: n indices without replacement, p ~ exp(beta * clearance / max clearance) (host side, numpy's
    generator like the reference, so a seed reproduces the reference's draw)."""
    if seed is not None:
        np.random.seed(seed)
    clearances = np.asarray(clearances)
    weight = np.exp(beta * clearances / clearances.max())
    return np.random.choice(range(len(weight)), size=n, p=weight / weight.sum(), replace=False)


class ImageSet(OrderedDict):
    """DataLoader.py:33-48: the assets of one imageset (name, lr, hr, hr_map, clearances)."""

    def __repr__(self):
        lines = [f"{'name':>10} : {self['name']}"]
        for key, v in self.items():
            lines.append(f"{key:>10} : {tuple(v.shape)} {type(v).__name__} ({v.dtype})" if hasattr(v, "shape")
                         else f"{key:>10} : {type(v).__name__} ({v})")
        return "\n".join(lines)


def _pick_views(imset_dir, top_k, beta, seed):
    """View ids of an imageset in the order the reference reads them (DataLoader.py:107-131) and their clearances."""
    ids = _view_ids(imset_dir)
    if not isfile(join(imset_dir, "clearance.npy")):
        raise Exception("please call the save_clearance.py before call DataLoader")
    clearances = np.load(join(imset_dir, "clearance.npy"))
    if top_k is not None and top_k > 0:
        picked = sample_clearest(clearances, n=min(top_k, len(ids)), beta=beta, seed=seed)
    else:
        picked = clearance_order(clearances)                                   # max to min
    return ids[picked], clearances[picked]


def read_imagesets(imset_dirs, seed=None, top_k=None, beta=0., threads=0):
    """Many imagesets at once: every LR view of every directory is decoded by ONE native call (one thread pool over
    hundreds of files instead of a pool per imageset), the HR images and status maps by a second and third one.  All
    imagesets must share one LR size.  Returns a list of ImageSets like read_imageset (no patch sampling)."""
    imset_dirs = list(imset_dirs)
    picked = [_pick_views(d, top_k, beta, seed) for d in imset_dirs]
    lr_paths = [join(d, f"LR{i}.png") for d, (ids, _) in zip(imset_dirs, picked) for i in ids]
    lrs = read_png_u16(lr_paths, threads=threads)
    maps = read_png_u16([join(d, "SM.png") for d in imset_dirs], threads=threads, pin=False).numpy().astype(bool)
    with_hr = [k for k, d in enumerate(imset_dirs) if exists(join(d, "HR.png"))]
    hrs = read_png_u16([join(imset_dirs[k], "HR.png") for k in with_hr], threads=threads) if with_hr else None
    out, at = [], 0
    for k, (d, (ids, clearances)) in enumerate(zip(imset_dirs, picked)):
        hr = hrs[with_hr.index(k)] if k in with_hr else None
        out.append(ImageSet(name=basename(d), lr=lrs[at:at + len(ids)], hr=hr, hr_map=maps[k], clearances=clearances))
        at += len(ids)
    return out


def read_imageset(imset_dir, create_patches=False, patch_size=64, seed=None, top_k=None, beta=0., threads=0):
    """DataLoader.py:73-148.  Returns an ImageSet whose 'lr' (L, H, W) and 'hr' are UINT16 tensors in pinned memory (the
    reference returns uint16 numpy arrays at this stage as well), 'hr_map' a bool array, 'clearances' the sorted scores."""
    ids, clearances = _pick_views(imset_dir, top_k, beta, seed)
    lr = read_png_u16([join(imset_dir, f"LR{i}.png") for i in ids], threads=threads)
    hr_map = read_png_u16([join(imset_dir, "SM.png")], threads=1, pin=False)[0].numpy().astype(bool)
    hr = read_png_u16([join(imset_dir, "HR.png")], threads=1)[0] if exists(join(imset_dir, "HR.png")) else None
    if create_patches:
        if seed is not None:
            np.random.seed(seed)
        x = np.random.randint(low=0, high=lr.shape[1] - patch_size)
        y = np.random.randint(low=0, high=lr.shape[2] - patch_size)
        lr = get_patch(lr, x, y, patch_size).contiguous()
        hr_map = get_patch(hr_map, x * 3, y * 3, patch_size * 3)
        if hr is not None:
            hr = get_patch(hr, x * 3, y * 3, patch_size * 3).contiguous()
    return ImageSet(name=basename(imset_dir), lr=lr, hr=hr, hr_map=hr_map, clearances=clearances)


class ImagesetDataset(torch.utils.data.Dataset):
    """DataLoader.py:153-205: imagesets from a list of directories; index by int, name or slice.  ``raw16=False`` (default)
    yields the reference's tensors ('lr' / 'hr' float32 in [0, 1], 'hr_map' float32); ``raw16=True`` keeps 'lr' / 'hr' as
    uint16 (pinned) for the device-side collate (predict.collate_device) -- same values after the on-device scaling."""

    def __init__(self, imset_dir, config, seed=None, top_k=-1, beta=0., raw16=False, threads=0):
        super().__init__()
        self.imset_dir = imset_dir
        self.name_to_dir = {basename(d): d for d in imset_dir}
        self.create_patches = config["create_patches"]
        self.patch_size = config["patch_size"]
        self.seed, self.top_k, self.beta, self.raw16, self.threads = seed, top_k, beta, raw16, threads

    def __len__(self):
        return len(self.imset_dir)

    def _load(self, directory):
        imset = read_imageset(directory, create_patches=self.create_patches, patch_size=self.patch_size, seed=self.seed,
                              top_k=self.top_k, beta=self.beta, threads=self.threads)
        return self._finish(imset)

    def _finish(self, imset):
        if not self.raw16:                                              # DataLoader.py:195-198
            imset["lr"] = (imset["lr"].to(torch.float32) / 65535.0)
            if imset["hr"] is not None:
                imset["hr"] = (imset["hr"].to(torch.float32) / 65535.0)
        if imset["hr"] is not None:
            imset["hr_map"] = torch.from_numpy(imset["hr_map"].astype(np.float32))
        return imset

    def __getitem__(self, index):
        if isinstance(index, int):
            dirs = [self.imset_dir[index]]
        elif isinstance(index, str):
            dirs = [self.name_to_dir[index]]
        elif isinstance(index, slice):
            dirs = self.imset_dir[index]
        else:
            raise KeyError("index must be int, string, or slice")
        if len(dirs) > 1 and not self.create_patches:          # a slice: one native decode call for all views of all imagesets
            try:
                imsets = [self._finish(im) for im in read_imagesets(dirs, seed=self.seed, top_k=self.top_k, beta=self.beta,
                                                                    threads=self.threads)]
            except RuntimeError:                               # imagesets of different sizes: one by one
                imsets = [self._load(d) for d in dirs]
        else:
            imsets = [self._load(d) for d in dirs]
        return imsets[0] if len(imsets) == 1 else imsets


def generate_submission_file(model, imset_dataset, out="../submission", min_L=16, batch_size=32, threads=0):
    """predict.py:161-194: one ``<imageset name>.png`` (16-bit) per imageset under ``out`` and ``out/submission.zip`` with
    every ``imgset*`` file.  Imagesets of equal shape go through the model ``batch_size`` at a time; the SR image is
    converted with ``img_as_uint`` ON THE DEVICE and crosses PCIe as uint16; PNG encoding runs on the native thread pool."""
    from .predict import collate_device, img_as_uint_u16
    os.makedirs(out, exist_ok=True)
    device = next(model.parameters()).device
    pending = {}

    def flush(group):
        lrs, alphas, _, _, names = collate_device(group, min_L, device)
        sr16 = img_as_uint_u16(model(lrs, alphas)[:, 0]).cpu()
        write_png_u16([os.path.join(out, name + ".png") for name in names], sr16, threads=threads)

    for imset in imset_dataset:
        group = pending.setdefault(tuple(imset["lr"].shape[1:]), [])
        group.append(imset)
        if len(group) == batch_size:
            flush(group)
            group.clear()
    for group in pending.values():
        if group:
            flush(group)
    members = sorted(f for f in os.listdir(out) if f.startswith("imgset"))       # predict.py:189: skip the archive itself
    archive = out + "/submission.zip"
    _lib.check(_lib.load().hrn_zip_store(os.fsencode(archive), _c_paths([os.path.join(out, f) for f in members]),
                                         _c_paths(members), len(members)), "hrn_zip_store")
    return archive
