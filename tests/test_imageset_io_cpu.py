"""CPU suite for the I/O ends of the path and the host half of the collate (N2): the native PNG
codec / clearance order / stored ZIP against the PIL-written fixture under tests/golden/imgsets/ (made by
tests/golden/make_png_fixtures.py) and against Python's own zipfile / zlib.  No GPU needed: these entry points are host code."""
import os
import zipfile

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FIX = os.path.join(ROOT, "tests", "golden", "imgsets")


@pytest.fixture(scope="module")
def io():
    from highres_net_b200 import imageset_io
    return imageset_io


@pytest.fixture(scope="module")
def arrays():
    return np.load(os.path.join(ROOT, "tests", "golden", "imgsets.npz"))


def _dirs():
    return [os.path.join(FIX, "RED", "imgset0001"), os.path.join(FIX, "RED", "imgset0002"), os.path.join(FIX, "NIR", "imgset0003")]


def test_png_decode_matches_the_fixture_pixels(io, arrays):
    for d in _dirs():
        name = os.path.basename(d)
        for f in sorted(os.listdir(d)):
            if not f.endswith(".png"):
                continue
            want = arrays[f"{name}/{f[:-4]}"]
            h, w, depth, ctype = io.png_info(os.path.join(d, f))
            assert (h, w) == want.shape and ctype == 0 and depth == (16 if want.dtype == np.uint16 else 8)
            got = io.read_png_u16([os.path.join(d, f)], pin=False)[0].numpy()
            assert got.dtype == np.uint16 and np.array_equal(got, want.astype(np.uint16)), f
    bits = io.read_png_u16([os.path.join(FIX, "bits.png")], pin=False)[0].numpy()
    assert io.png_info(os.path.join(FIX, "bits.png"))[2] == 1 and np.array_equal(bits, arrays["bits"].astype(np.uint16))


def test_png_decode_many_files_on_the_thread_pool(io, arrays):
    d = _dirs()[0]
    paths = [os.path.join(d, f"LR{v:03d}.png") for v in range(5)] * 7
    for threads in (1, 3, 0):
        got = io.read_png_u16(paths, threads=threads, pin=False).numpy()
        for k, p in enumerate(paths):
            assert np.array_equal(got[k], arrays["imgset0001/" + os.path.basename(p)[:-4]])


def test_png_decode_agrees_with_pil(io):
    PIL = pytest.importorskip("PIL.Image")
    for d in _dirs():
        for f in sorted(os.listdir(d)):
            if f.endswith(".png"):
                assert np.array_equal(io.read_png_u16([os.path.join(d, f)], pin=False)[0].numpy(),
                                      np.array(PIL.open(os.path.join(d, f))).astype(np.uint16))


def test_png_errors_are_loud(io, tmp_path):
    with pytest.raises(RuntimeError, match="cannot open"):
        io.png_info(str(tmp_path / "missing.png"))
    bad = tmp_path / "bad.png"
    bad.write_bytes(b"not a png at all, just bytes" * 4)
    with pytest.raises(RuntimeError, match="not a PNG"):
        io.read_png_u16([str(bad)], pin=False)
    good = open(os.path.join(_dirs()[0], "LR000.png"), "rb").read()
    flipped = tmp_path / "crc.png"
    flipped.write_bytes(good[:60] + bytes([good[60] ^ 0xff]) + good[61:])
    with pytest.raises(RuntimeError, match="CRC|zlib"):
        io.read_png_u16([str(flipped)], pin=False)
    with pytest.raises(RuntimeError, match="expected"):       # second file has another size
        io.read_png_u16([os.path.join(_dirs()[0], "LR000.png"), os.path.join(_dirs()[0], "SM.png")], pin=False)


def test_png_encode_round_trip_and_pil_reads_it(io, tmp_path):
    rng = np.random.RandomState(3)
    imgs = np.concatenate([(rng.rand(2, 40, 56) * 65535).astype(np.uint16),
                           np.full((1, 40, 56), 65535, np.uint16), np.zeros((1, 40, 56), np.uint16)])
    paths = [str(tmp_path / f"imgset{i:04d}.png") for i in range(len(imgs))]
    io.write_png_u16(paths, imgs, threads=2)
    assert np.array_equal(io.read_png_u16(paths, pin=False).numpy(), imgs)
    PIL = pytest.importorskip("PIL.Image")
    for p, want in zip(paths, imgs):
        back = np.array(PIL.open(p))
        assert back.dtype == np.uint16 and np.array_equal(back, want)
    with pytest.raises(TypeError):
        io.write_png_u16(paths[:1], imgs[:1].astype(np.float32))


def test_clearance_scores_and_order(io, arrays, tmp_path):
    import shutil
    work = tmp_path / "imgset0001"
    shutil.copytree(_dirs()[0], work)
    io.save_clearance_scores([str(work)])
    got = np.load(work / "clearance.npy")
    want = np.array([arrays[f"imgset0001/QM{v:03d}"].astype(np.uint16).sum() for v in range(5)])    # save_clearance.py:24
    assert np.array_equal(got, want)
    order = io.clearance_order(got)
    assert np.array_equal(order, np.argsort(got, kind="stable")[::-1])      # DataLoader.py:128 with the tie rule stated
    assert list(order[:2]) == [3, 1]                                        # the two fully clear views, higher index first
    tie_free = np.array([3.0, 9.0, 1.0, 7.5, 8.25] * 5) + np.arange(25) * 1e-3
    assert np.array_equal(io.clearance_order(tie_free), np.argsort(tie_free)[::-1])


def test_read_imageset_and_dataset_mirror_the_reference_loader(io, arrays, tmp_path):
    import shutil
    dirs = []
    for d in _dirs():
        dst = tmp_path / os.path.basename(d)
        shutil.copytree(d, dst)
        dirs.append(str(dst))
    with pytest.raises(Exception, match="save_clearance"):                  # DataLoader.py:117-119
        io.read_imageset(dirs[0])
    io.save_clearance_scores(dirs)
    imset = io.read_imageset(dirs[0])
    clr = np.load(os.path.join(dirs[0], "clearance.npy"))
    order = np.argsort(clr, kind="stable")[::-1]
    assert imset["name"] == "imgset0001" and imset["lr"].dtype == torch.uint16 and tuple(imset["lr"].shape) == (5, 24, 24)
    for k, v in enumerate(order):
        assert np.array_equal(imset["lr"][k].numpy(), arrays[f"imgset0001/LR{v:03d}"])
    assert np.array_equal(imset["clearances"], clr[order])
    assert imset["hr_map"].dtype == bool and np.array_equal(imset["hr_map"], arrays["imgset0001/SM"] > 0)
    assert np.array_equal(imset["hr"].numpy(), arrays["imgset0001/HR"])
    assert io.read_imageset(dirs[2])["hr"] is None                          # test split: no HR.png
    # the Dataset: float32 tensors in [0, 1] like DataLoader.py:195-198 (x / 65535 rounded once to fp32)
    ds = io.ImagesetDataset(dirs, {"create_patches": False, "patch_size": 8})
    item = ds[0]
    want = (arrays[f"imgset0001/LR{order[0]:03d}"].astype(np.float64) / 65535).astype(np.float32)
    assert item["lr"].dtype == torch.float32 and np.array_equal(item["lr"][0].numpy(), want)
    assert item["hr_map"].dtype == torch.float32 and ds["imgset0002"]["name"] == "imgset0002" and len(ds[0:2]) == 2
    # a slice goes through ONE native decode call for all views of all imagesets (read_imagesets): same content
    both = io.ImagesetDataset(dirs, {"create_patches": False, "patch_size": 8}, raw16=True)
    for one, many in zip([both[0], both[1], both[2]], both[0:3]):
        assert one["name"] == many["name"] and torch.equal(one["lr"], many["lr"]) and np.array_equal(one["clearances"], many["clearances"])
        assert (one["hr"] is None) == (many["hr"] is None) and (one["hr"] is None or torch.equal(one["hr"], many["hr"]))
        assert np.array_equal(np.asarray(one["hr_map"]), np.asarray(many["hr_map"]))
    raw = io.ImagesetDataset(dirs, {"create_patches": False, "patch_size": 8}, raw16=True)[1]
    assert raw["lr"].dtype == torch.uint16 and tuple(raw["lr"].shape) == (3, 24, 24)
    # sampled / patched variant: same draws as the reference for a seed (np.random.seed -> choice / randint)
    sampled = io.read_imageset(dirs[0], create_patches=True, patch_size=8, seed=5, top_k=3, beta=50.0)
    np.random.seed(5)
    e_c = np.exp(50.0 * clr / clr.max())
    pick = np.random.choice(range(5), size=3, p=e_c / e_c.sum(), replace=False)
    np.random.seed(5)
    x, y = np.random.randint(0, 16), np.random.randint(0, 16)
    assert tuple(sampled["lr"].shape) == (3, 8, 8) and sampled["hr_map"].shape == (24, 24)
    assert np.array_equal(sampled["lr"][0].numpy(), arrays[f"imgset0001/LR{pick[0]:03d}"][x:x + 8, y:y + 8])


def test_zip_store_is_a_plain_stored_archive(io, tmp_path):
    import ctypes
    from highres_net_b200 import _lib
    files = []
    for i in range(3):
        p = tmp_path / f"imgset{i:04d}.png"
        p.write_bytes(os.urandom(1000 + 17 * i))
        files.append(str(p))
    arc = str(tmp_path / "submission.zip")
    _lib.check(_lib.load().hrn_zip_store(arc.encode(), io._c_paths(files), io._c_paths([os.path.basename(f) for f in files]), 3), "zip")
    with zipfile.ZipFile(arc) as z:
        assert z.testzip() is None and z.namelist() == [os.path.basename(f) for f in files]
        assert all(info.compress_type == zipfile.ZIP_STORED for info in z.infolist())      # ZipFile(mode='w') default, predict.py:187
        for f in files:
            assert z.read(os.path.basename(f)) == open(f, "rb").read()


def test_host_collate_matches_the_oracle_restatement(io):
    """utils.collateFunction semantics (utils.py:63-113): truncate at min_L, zero-pad with alpha 0, the hr rule."""
    from highres_net_b200.predict import collateFunction
    from oracle import predict_oracle
    rng = np.random.RandomState(1)
    batch = [{"name": f"s{i}", "lr": torch.from_numpy(rng.rand(n, 6, 6).astype(np.float32)),
              "hr": torch.from_numpy(rng.rand(18, 18).astype(np.float32)),
              "hr_map": torch.from_numpy((rng.rand(18, 18) > 0.2).astype(np.float32))} for i, n in enumerate((3, 9, 5, 1))]
    for min_l in (1, 4, 5, 16):
        lrs, alphas, hrs, hms, names = collateFunction(min_L=min_l)(batch)
        ref = predict_oracle.collate(batch, min_l)
        assert torch.equal(lrs, torch.as_tensor(ref[0])) and torch.equal(alphas, torch.as_tensor(ref[1]))
        assert torch.equal(hrs, torch.stack([b["hr"] for b in batch])) and torch.equal(hms, torch.stack([b["hr_map"] for b in batch]))
        assert names == [b["name"] for b in batch]
    batch[2]["hr"] = None                                                    # test-set imageset in the batch
    lrs, alphas, hrs, hms, names = collateFunction(min_L=4)(batch)
    assert isinstance(hrs, list) and len(hrs) == 2 and isinstance(hms, list) and len(hms) == 4
    raw = [{"name": "u", "lr": torch.from_numpy((rng.rand(2, 6, 6) * 65535).astype(np.uint16)), "hr": None, "hr_map": None}]
    lrs, alphas, _, _, _ = collateFunction(min_L=3)(raw)
    assert lrs.dtype == torch.float32 and torch.equal(lrs[0, :2], raw[0]["lr"].to(torch.float32) / 65535.0)
    assert alphas.tolist() == [[1.0, 1.0, 0.0]]
