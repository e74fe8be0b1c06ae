"""CPU suite: the oracle restatement against the fixtures minted from the
unmodified reference (oracle/make_golden.py).  This is what pins the oracle."""
import numpy as np
import torch
import pytest

from oracle import cases, hrnet_oracle, scoring_oracle


def test_param_inventory_matches_reference():
    shapes = hrnet_oracle.param_shapes()
    assert len(shapes) == 31
    assert sum(int(np.prod(s)) for s in shapes.values()) == 591818   # paper txt:824


@pytest.mark.parametrize("name", [n for n in cases.HRNET_CASES if n != "c1_b2_l4_s128"])
def test_hrnet_oracle_matches_reference_golden(golden, name):
    params = hrnet_oracle.make_params(cases.WEIGHT_SEED)
    lrs, alphas = cases.hrnet_inputs(name)
    sr = hrnet_oracle.hrnet_forward(params, lrs, alphas).numpy()
    ref = golden["hrnet_forward"][name]
    assert sr.shape == ref.shape
    # same ATen ops; thread-count dependent summation only (1.1e-7)
    assert np.abs(sr - ref).max() <= 2e-6


def test_hrnet_oracle_c1_config(golden):
    """BASELINE.json configs[0]: B2 L4 128x128 fp32 on CPU."""
    params = hrnet_oracle.make_params(cases.WEIGHT_SEED)
    lrs, alphas = cases.hrnet_inputs("c1_b2_l4_s128")
    sr = hrnet_oracle.hrnet_forward(params, lrs, alphas).numpy()
    assert sr.shape == (2, 1, 384, 384)
    assert np.abs(sr - golden["hrnet_forward"]["c1_b2_l4_s128"]).max() <= 2e-6


def test_hrnet_oracle_fp64_arbiter(golden):
    params = hrnet_oracle.make_params(cases.WEIGHT_SEED)
    lrs, alphas = cases.hrnet_inputs("b2_l4_s32")
    import torch
    sr64 = hrnet_oracle.hrnet_forward(params, lrs, alphas, dtype=torch.float64).numpy()
    assert np.abs(sr64 - golden["hrnet_forward"]["b2_l4_s32__f64"]).max() <= 1e-12


def test_odd_view_is_dropped():
    """HRNet.py:115 - with odd L the last view never reaches the output."""
    params = hrnet_oracle.make_params(cases.WEIGHT_SEED)
    lrs, alphas = cases.hrnet_inputs("b1_l5_s24")
    a = hrnet_oracle.hrnet_forward(params, lrs, alphas).numpy()
    lrs2 = lrs.copy()
    # keep the median unchanged: view 4 only enters through the anchor, so perturb
    # it towards a value that leaves the lower median (index 2 of 5) in place
    enc = {}
    hrnet_oracle.hrnet_forward(params, lrs, alphas, trace=enc)
    fused_a = enc["fused"].numpy()
    feats = enc["encoded"].clone()
    feats[:, 4] += 1.0
    import torch
    fused_b = hrnet_oracle.fuse(params, feats, torch.from_numpy(alphas)).numpy()
    assert np.array_equal(fused_a, fused_b)
    assert a.shape == (1, 1, 72, 72) and lrs2.shape == lrs.shape


def test_alpha_zero_is_skip_connection():
    """alphas [1,1,0,0] -> result == h0 + fuse(cat(h0,h1))."""
    import torch
    params = hrnet_oracle.make_params(cases.WEIGHT_SEED)
    rng = np.random.RandomState(5)
    x = torch.from_numpy(rng.randn(1, 4, 64, 8, 8).astype(np.float32))
    al = torch.tensor([[1.0, 1.0, 0.0, 0.0]])
    lvl1, al1 = hrnet_oracle.fuse_level(params, x, al)
    assert torch.equal(lvl1, x[:, :2])          # both bobs are padded views
    out = hrnet_oracle.fuse(params, x, al)
    direct, _ = hrnet_oracle.fuse_level(params, x[:, :2], torch.ones(1, 2))
    assert torch.allclose(out, direct[:, 0], atol=1e-6)


@pytest.mark.parametrize("name", list(cases.LANCZOS_CASES))
def test_lanczos_oracle_matches_reference_golden(golden, name):
    img, shift, p = cases.lanczos_inputs(name)
    out = scoring_oracle.lanczos_shift(img, shift, p=p)
    assert np.abs(out - golden["lanczos"][name]).max() <= 2e-6


def test_lanczos_taps_known_answers(golden):
    taps = scoring_oracle.lanczos_taps(np.array(cases.LANCZOS_TAP_SHIFTS, dtype=np.float32))
    assert np.abs(taps - golden["lanczos"]["taps"]).max() <= 3e-7
    assert np.allclose(taps.sum(1), 1.0, atol=1e-6)                      # lanczos.py:41
    # SURVEY.md section 9 KAT: lanczos_kernel(0.3)
    kat = np.array([0.0070, 0.0310, -0.1419, 0.8417, 0.3348, -0.0830, 0.0104], dtype=np.float32)
    assert np.abs(taps[1] - kat).max() < 6e-5
    assert np.argmax(taps[0]) == 3 and np.argmax(taps[2]) == 4           # d=0 -> tap 3, d=1 -> tap 4


@pytest.mark.parametrize("name", list(cases.CPSNR_CASES))
def test_cpsnr_oracle_matches_reference_golden(golden, name):
    sr, hr, hm = cases.cpsnr_inputs(name)
    g = golden["cpsnr"]
    for i in range(sr.shape[0]):
        mx, am, sites = scoring_oracle.shift_cpsnr(sr[i], hr[i], hm[i])
        assert np.array_equal(np.float32(mx), g[name + "__max"][i], equal_nan=True)   # bit-exact, same numpy ops
        assert int(am) == int(g[name + "__argmax"][i])
        assert np.array_equal(sites.astype(np.float32), g[name + "__sites"][i], equal_nan=True)


def test_cpsnr_known_shift_convention():
    """sr = roll(hr, (+1, -2)) -> best site 19 = (x=2, y=5)."""
    rng = np.random.RandomState(11)
    hr = rng.rand(64, 64).astype(np.float32)
    sr = np.roll(hr, (1, -2), axis=(0, 1))
    mx, am, _ = scoring_oracle.shift_cpsnr(sr, hr, np.ones_like(hr))
    assert int(am) == 19 and np.isinf(mx)


@pytest.mark.parametrize("name", list(cases.PREDICT_CASES))
def test_predict_oracle_matches_reference_golden(golden, name):
    """predict.get_sr_and_score (predict.py:17-49) restated on the oracle pieces vs the reference run."""
    from oracle import predict_oracle
    g = golden["predict"]
    n, s, has_hr = cases.PREDICT_CASES[name]
    imset = {"lr": cases.predict_lrs(name), "hr": g[name + "__hr"] if has_hr else None,
             "hr_map": g[name + "__hr_map"] if has_hr else None}
    sr, score = predict_oracle.get_sr_and_score(imset, hrnet_oracle.make_params(cases.WEIGHT_SEED), cases.PREDICT_MIN_L)
    assert np.abs(sr - g[name + "__sr"]).max() <= 2e-6
    if has_hr:
        assert abs(float(score) - float(g[name + "__score"])) <= 1e-4
    else:
        assert score is None and np.isnan(g[name + "__score"])


def test_collate_mirror_matches_oracle():
    """Host logic of the product's collateFunction (utils.py:63-113 mirror) against the oracle restatement."""
    import torch
    from highres_net_b200.predict import collateFunction
    from oracle import predict_oracle
    rng = np.random.RandomState(9)
    imsets = [{"name": f"s{i}", "lr": torch.from_numpy(rng.rand(n, 8, 8).astype(np.float32)),
               "hr": torch.from_numpy(rng.rand(24, 24).astype(np.float32)), "hr_map": torch.ones(24, 24)}
              for i, n in enumerate((3, 16, 21))]
    lrs, alphas, hrs, hms, names = collateFunction(min_L=16)(imsets)
    o_lrs, o_alphas = predict_oracle.collate(imsets, 16)
    assert lrs.shape == (3, 16, 8, 8) and np.array_equal(lrs.numpy(), o_lrs) and np.array_equal(alphas.numpy(), o_alphas)
    assert hrs.shape == (3, 24, 24) and hms.shape == (3, 24, 24) and names == ["s0", "s1", "s2"]
    assert alphas[0].tolist() == [1.0] * 3 + [0.0] * 13 and float(lrs[0, 3:].abs().max()) == 0.0
    imsets[1]["hr"] = None                      # a test-split imageset: hr batch stays a (partial) list like the reference
    _, _, hrs2, hms2, _ = collateFunction(min_L=16)(imsets)
    assert isinstance(hrs2, list) and len(hrs2) == 1 and isinstance(hms2, list)


# ---------------------------------------------------------------------------- train.get_loss
@pytest.mark.parametrize("name", list(cases.LOSS_CASES))
@pytest.mark.parametrize("metric", cases.LOSS_METRICS)
def test_clear_loss_oracle_matches_reference_golden(golden, name, metric):
    sr, hr, hm = cases.loss_inputs(name)
    mine = scoring_oracle.clear_loss(sr, hr, hm, metric)
    ref = golden["loss"][f"{name}__{metric}"]
    assert mine.shape == ref.shape
    assert np.abs(mine / ref - 1).max() <= 2e-5            # fp32 summation order only


def test_crop_mask_oracle_matches_reference_golden(golden):
    for ps, cs in ((32, 3), (4, 1), (64, 6)):
        assert np.array_equal(scoring_oracle.crop_mask(ps, cs), golden["loss"][f"crop_{ps}_{cs}"])


@pytest.mark.parametrize("name", list(cases.APPLY_SHIFTS_CASES))
def test_apply_shifts_oracle_matches_reference_golden(golden, name):
    images, thetas = cases.apply_shifts_inputs(name)
    out = scoring_oracle.apply_shifts(images, thetas)
    assert np.abs(out - golden["apply_shifts"][name]).max() <= 2e-6


def test_evaluate_oracle_matches_reference_golden(golden):
    """predict.evaluate of the reference (predict.py:103-135) on the tiny datasets of cases.EVALUATE_SETS: the oracle's
    get_sr_and_score reproduces every score."""
    from oracle import predict_oracle
    g = golden["evaluate"]
    params = hrnet_oracle.make_params(cases.WEIGHT_SEED)
    for name, (case, split) in cases.EVALUATE_SETS.items():
        has_hr = cases.PREDICT_CASES[case][2]
        im = {"lr": cases.evaluate_lrs(name), "hr": g[name + "__hr"] if has_hr else None,
              "hr_map": g[name + "__hr_map"] if has_hr else None}
        _, score = predict_oracle.get_sr_and_score(im, params, cases.PREDICT_MIN_L)
        if has_hr:
            assert abs(float(score) - float(g[name + "__score"])) <= 1e-4
        else:
            assert score is None and np.isnan(g[name + "__score"])


def test_benchmark_mirror_matches_reference_golden(golden):
    """predict.benchmark (predict.py:138-158) is host-side pandas: same columns, index and values as the reference table."""
    import importlib
    predict = importlib.import_module("highres_net_b200.predict")
    g = golden["evaluate"]
    names = [str(n) for n in g["benchmark__index"]]
    scores = {n: np.float32(g[n + "__score"]) for n in names}
    part = {n: cases.EVALUATE_SETS[n][1] for n in names}
    clr = {n: cases.evaluate_clearances(n) for n in names}
    table = predict.benchmark({n: cases.EVALUATE_BASELINE[n] for n in names}, scores, part, clr)
    assert list(table.columns) == [str(c) for c in g["benchmark__columns"]]
    assert list(table.index) == names and list(table["part"]) == [str(p) for p in g["benchmark__part"]]
    for col in ("ESA", "model", "score", "mean_clr", "std_clr"):
        assert np.allclose(table[col].to_numpy(dtype=np.float64), g["benchmark__" + col], rtol=1e-6, atol=0), col


def test_shiftnet_oracle_matches_reference_golden(golden):
    """ShiftNet.forward in eval mode (ShiftNet.py:49-75) and train.register_batch (train.py:26-44) of the reference."""
    from oracle import shiftnet_oracle
    g = golden["shiftnet"]
    params = shiftnet_oracle.make_params(0)
    theta = shiftnet_oracle.shiftnet_forward(params, shiftnet_oracle.make_pairs(6, 0)).numpy()
    assert np.abs(theta - g["theta"]).max() <= 1e-5
    assert np.abs(g["theta"] - g["theta_fp64"]).max() <= 1e-5                     # fp64 arbiter of the reference itself
    pairs = shiftnet_oracle.make_pairs(6, 1).reshape(2, 3, 2, 128, 128)
    thetas = shiftnet_oracle.register_batch(params, pairs[:, :, 1], pairs[:, 0, 0][:, None]).numpy()
    assert thetas.shape == (2, 3, 2) and np.abs(thetas - g["register_thetas"]).max() <= 1e-5


def test_shiftnet_mirror_has_the_reference_state_dict(golden):
    """Same keys in the same order as the reference module, 34,187,648 parameters; a
    checkpoint of the reference loads with strict=True.  Construction needs no GPU; forward does."""
    import importlib
    from oracle import shiftnet_oracle
    shiftnet = importlib.import_module("highres_net_b200.shiftnet")
    g = golden["shiftnet"]
    net = shiftnet.ShiftNet()
    assert list(net.state_dict().keys()) == [str(k) for k in g["state_dict_keys"]]
    assert sum(p.numel() for p in net.parameters()) == int(g["n_params"]) == 34187648
    assert float(net.fc2.weight.abs().max()) == 0.0                               # ShiftNet.py:48
    net.load_state_dict(shiftnet_oracle.make_params(0), strict=True)
    with pytest.raises(RuntimeError):
        net.eval()(torch.zeros(1, 2, 128, 128))                                   # CPU tensor: no fallback
    with pytest.raises(ValueError):
        shiftnet.ShiftNet(in_channel=2)


def test_trainstep_forward_oracle_matches_reference_golden(golden):
    """Forward value of the reference training step (train.py:174-187: HRNet -> register_batch on the centre crops ->
    apply_shifts -> -cPSNR loss with the crop mask + lambda * mean(shift)^2), composed from the oracle pieces."""
    from oracle import shiftnet_oracle
    from oracle.make_golden_trainstep import inputs
    g = golden["trainstep"]
    lrs, alphas = inputs()
    off = int(g["offset"])
    srs = hrnet_oracle.hrnet_forward(hrnet_oracle.make_params(cases.WEIGHT_SEED), lrs, alphas).numpy()
    assert np.abs(srs - g["srs"]).max() <= 2e-6
    shifts = shiftnet_oracle.register_batch(shiftnet_oracle.make_params(0), srs[:, :, off:off + 128, off:off + 128],
                                            g["hr"][:, off:off + 128, off:off + 128][:, None]).numpy()
    assert np.abs(shifts - g["shifts"]).max() <= 1e-5
    shifted = scoring_oracle.apply_shifts(srs, shifts)[:, 0]
    assert np.abs(shifted - g["srs_shifted"]).max() <= 1e-5
    mask = scoring_oracle.crop_mask(64, int(g["crop"]))[0] * g["hr_map"]
    loss = -scoring_oracle.clear_loss(shifted, g["hr"], mask, "cPSNR")
    assert np.abs(loss - g["loss"]).max() <= 1e-3
    total = loss.mean() + float(g["lam"]) * shifts.mean() ** 2
    assert abs(total - float(g["total"])) <= 1e-3
