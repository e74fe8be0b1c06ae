"""GPU parity tests (run on the B200 box: pytest -m gpu).  Every check goes through the
C ABI (ctypes -> libhrn_b200.so) and compares against the CPU oracle and/or the golden
fixtures minted from the unmodified reference.  Nothing here reads /root/reference.

Gates (BASELINE.json north_star): SR max-abs <= 1e-2 vs the fp32 reference (bf16
activations, fp32 tensor-core accumulation); cPSNR within 0.01 dB; best shift bit-exact;
Lanczos <= 1e-5 max-abs (fp32 kernel).  Tighter regression gates sit beside them so a
real bug cannot hide inside the loose north-star tolerance."""
import numpy as np
import pytest
import torch

from oracle import cases, hrnet_oracle, scoring_oracle

pytestmark = pytest.mark.gpu

SR_GATE = 1e-2            # north star
SR_REGRESSION_GATE = 1.5e-3
LAYER_REL_GATE = 1.5e-2   # per-stage error relative to the stage's max magnitude (bf16 storage)
CPSNR_GATE_DB = 0.01      # north star
CPSNR_KERNEL_GATE_DB = 1e-4
LANCZOS_GATE = 1e-5


@pytest.fixture(scope="module")
def hb():
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device; there is no CPU fallback")
    import highres_net_b200 as m
    import os
    assert os.path.exists(m.library_path()), "libhrn_b200.so missing"
    return m


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def net(hb, dev):
    model = hb.HRNet(hrnet_oracle.DEFAULT_NETWORK_CONFIG).eval()
    model.load_state_dict(hrnet_oracle.make_params(cases.WEIGHT_SEED))
    return model.to(dev)


# ---------------------------------------------------------------------------- HRNet.forward
@pytest.mark.parametrize("name", list(cases.HRNET_CASES))
def test_hrnet_forward_matches_reference_golden(hb, net, dev, golden, name):
    lrs, alphas = cases.hrnet_inputs(name)
    before = hb.kernel_launch_count()
    sr = net(torch.from_numpy(lrs).to(dev), torch.from_numpy(alphas).to(dev))
    assert hb.kernel_launch_count() > before, "forward did not launch any native kernel"
    ref = golden["hrnet_forward"][name]
    assert sr.dtype == torch.float32 and tuple(sr.shape) == ref.shape and sr.device.type == "cuda"
    err = np.abs(sr.cpu().numpy() - ref).max()
    assert err <= SR_GATE
    assert err <= SR_REGRESSION_GATE, err


@pytest.mark.parametrize("name", ["b2_l4_s32", "b2_l9_s16", "b1_l16_s16", "b1_l5_s24"])
def test_hrnet_stages_match_oracle(net, dev, name):
    from highres_net_b200 import hrnet as hm
    params = hrnet_oracle.make_params(cases.WEIGHT_SEED)
    lrs, alphas = cases.hrnet_inputs(name)
    b, l, s, _ = lrs.shape
    tr = {}
    hrnet_oracle.hrnet_forward(params, lrs, alphas, trace=tr)
    tl, ta = torch.from_numpy(lrs).to(dev), torch.from_numpy(alphas).to(dev)
    _, anchor = net.forward_stage(tl, ta, hm.stage_anchor(), (b, 1, s, s))
    assert torch.equal(anchor.cpu()[:, 0], tr["anchor"])                 # lower median: exact selection
    _, enc = net.forward_stage(tl, ta, hm.stage_enc(5), (b * l, 64, s, s))
    ref = tr["encoded"].reshape(b * l, 64, s, s)
    assert (enc.cpu() - ref).abs().max() <= LAYER_REL_GATE * ref.abs().max()
    n, level = l, 0
    while n // 2 > 0:
        half = n // 2
        _, lv = net.forward_stage(tl, ta, hm.stage_fuse(level, 2), (b * half, 64, s, s))
        ref = tr["levels"][level].reshape(b * half, 64, s, s)
        assert (lv.cpu() - ref).abs().max() <= LAYER_REL_GATE * ref.abs().max(), (name, level)
        n, level = half, level + 1


def _conv_layer_ref(x, w, b, slope):
    import torch.nn.functional as F
    y = F.conv2d(x, w.to(torch.bfloat16).to(torch.float32), b, padding=1)
    return F.prelu(y, slope)


@pytest.mark.parametrize("taps", [[(1, 0)], [(1, 2)], [(0, 1), (2, 2)], None])
def test_conv64_single_taps(hb, dev, taps):
    """tcgen05 conv in isolation: keep only some of the 9 taps (each tap = one shifted smem descriptor)."""
    from highres_net_b200 import hrnet as hm
    params = dict(hrnet_oracle.make_params(0))
    key = "encode.res_layers.0.block.0"
    w = params[key + ".weight"].clone()
    if taps is not None:
        mask = torch.zeros(3, 3)
        for ky, kx in taps:
            mask[ky, kx] = 1
        w = w * mask
    params[key + ".weight"] = w
    model = hb.HRNet(hrnet_oracle.DEFAULT_NETWORK_CONFIG).eval()
    model.load_state_dict(params)
    model = model.to(dev)
    s = 40
    lrs = torch.from_numpy(np.random.RandomState(7).rand(2, 1, s, s).astype(np.float32)).to(dev)
    al = torch.ones(2, 1, device=dev)
    _, x0 = model.forward_stage(lrs, al, hm.stage_enc(0), (2, 64, s, s))
    _, y1 = model.forward_stage(lrs, al, hm.stage_enc(1), (2, 64, s, s))
    ref = _conv_layer_ref(x0.cpu(), w, params[key + ".bias"], params["encode.res_layers.0.block.1.weight"])
    # same bf16 inputs and weights on both sides: only accumulation order and the bf16 store differ
    assert (y1.cpu() - ref).abs().max() <= 4e-3 * max(1.0, float(ref.abs().max()))


@pytest.mark.parametrize("b,l,size", [(1, 2, 1), (2, 3, 2), (1, 4, 3), (3, 7, 5), (1, 32, 8), (2, 2, 130), (1, 3, 257)])
def test_hrnet_edge_shapes_match_oracle(net, dev, b, l, size):
    """Tiny images (a strip of 1-3 rows, one ragged 128-pixel tile), deep recursion (L = 32: five fusion levels) and
    widths just past one / two column tiles, against the oracle computed on the fly."""
    rng = np.random.RandomState(100 * b + 10 * l + size)
    lrs = rng.rand(b, l, size, size).astype(np.float32)
    alphas = np.ones((b, l), dtype=np.float32)
    if l > 2:
        lrs[0, -1] = 0.0
        alphas[0, -1] = 0.0
    ref = hrnet_oracle.hrnet_forward(hrnet_oracle.make_params(cases.WEIGHT_SEED), lrs, alphas).numpy()
    sr = net(torch.from_numpy(lrs).to(dev), torch.from_numpy(alphas).to(dev))
    assert tuple(sr.shape) == (b, 1, 3 * size, 3 * size)
    err = np.abs(sr.cpu().numpy() - ref).max()
    assert err <= SR_GATE and err <= 2e-3, err          # L = 32 accumulates five levels of bf16 rounding: ~1.1e-3


@pytest.mark.parametrize("num_layers,alpha_residual", [(0, True), (1, True), (3, True), (2, False)])
def test_config_variants_match_oracle(hb, dev, num_layers, alpha_residual):
    """config.json knobs the kernels are not specialised away from: encoder depth and recursive.alpha_residual."""
    import copy
    cfg = copy.deepcopy(hrnet_oracle.DEFAULT_NETWORK_CONFIG)
    cfg["encoder"]["num_layers"] = num_layers
    cfg["recursive"]["alpha_residual"] = alpha_residual
    params = hrnet_oracle.make_params(5, cfg)
    model = hb.HRNet(cfg).eval()
    model.load_state_dict(params)
    model = model.to(dev)
    rng = np.random.RandomState(31 + num_layers)
    lrs = rng.rand(2, 4, 24, 24).astype(np.float32)
    alphas = np.array([[1, 1, 1, 1], [1, 1, 1, 0]], dtype=np.float32)
    lrs[1, 3] = 0.0
    ref = hrnet_oracle.hrnet_forward(params, lrs, alphas, cfg).numpy()
    sr = model(torch.from_numpy(lrs).to(dev), torch.from_numpy(alphas).to(dev)).cpu().numpy()
    assert np.abs(sr - ref).max() <= SR_REGRESSION_GATE


def test_unsupported_config_is_rejected(hb, dev):
    import copy
    cfg = copy.deepcopy(hrnet_oracle.DEFAULT_NETWORK_CONFIG)
    cfg["encoder"]["channel_size"] = 32
    cfg["recursive"]["in_channels"] = 32
    model = hb.HRNet(cfg).eval().to(dev)
    with pytest.raises(RuntimeError, match="unsupported network config"):
        model(torch.rand(1, 2, 16, 16, device=dev), torch.ones(1, 2, device=dev))


def test_alpha_zero_views_are_skipped(net, dev):
    """utils.py:89-95 contract: a padded view (alpha = 0) must not change the fused state (HRNet.py:123-128)."""
    rng = np.random.RandomState(3)
    lrs = rng.rand(1, 4, 32, 32).astype(np.float32)
    lrs[:, 2:] = 0.0
    alphas = np.array([[1, 1, 0, 0]], dtype=np.float32)
    params = hrnet_oracle.make_params(cases.WEIGHT_SEED)
    sr = net(torch.from_numpy(lrs).to(dev), torch.from_numpy(alphas).to(dev)).cpu().numpy()
    ref = hrnet_oracle.hrnet_forward(params, lrs, alphas).numpy()
    assert np.abs(sr - ref).max() <= SR_REGRESSION_GATE
    alphas_on = np.ones_like(alphas)
    sr_on = net(torch.from_numpy(lrs).to(dev), torch.from_numpy(alphas_on).to(dev)).cpu().numpy()
    assert np.abs(sr_on - sr).max() > 1e-4         # the alpha mask really is consulted


@pytest.mark.parametrize("l,pattern", [
    (4, [[1, 0, 1, 1], [0, 1, 1, 0]]),                       # holes: a dead alice next to a live bob must still be right
    (8, [[1, 1, 1, 0, 0, 1, 0, 1], [1, 0, 0, 0, 0, 0, 0, 0]]),
    (7, [[1, 1, 1, 1, 1, 0, 0], [0, 0, 0, 0, 0, 0, 1]]),     # odd L (last view dropped) and alpha[0] = 0
    (16, [[1] * 9 + [0] * 7, [1] * 16]),                     # Proba-V style trailing padding
    (3, [[0, 0, 0], [1, 0, 1]]),
])
def test_dead_view_skipping_is_exact(hb, dev, l, pattern):
    """Views / pairs that cannot reach the output are not computed (live-work lists).  Whatever the alpha pattern, the
    result must equal the oracle (which computes everything and multiplies by alpha, HRNet.py:123-128) and must be
    bit-identical to the dense run of the same kernels."""
    params = hrnet_oracle.make_params(cases.WEIGHT_SEED)
    model = hb.HRNet(hrnet_oracle.DEFAULT_NETWORK_CONFIG).eval()
    model.load_state_dict(params)
    model = model.to(dev)
    rng = np.random.RandomState(l)
    lrs = rng.rand(len(pattern), l, 24, 24).astype(np.float32)
    alphas = np.array(pattern, dtype=np.float32)
    tl, ta = torch.from_numpy(lrs).to(dev), torch.from_numpy(alphas).to(dev)
    ref = hrnet_oracle.hrnet_forward(params, lrs, alphas).numpy()
    for wave in (1, 0):                     # wavefront schedule (out of place: carried views) and three launches (in place)
        model.debug_set(dev, "fuse_wave", wave)
        model.debug_set(dev, "skip_dead_views", 1)
        # poison the workspace first, so that a skipped view that is read anyway shows up
        model(torch.full_like(tl, 1e4), torch.ones_like(ta))
        sr = model(tl, ta).cpu().numpy()
        assert np.isfinite(sr).all()
        assert np.abs(sr - ref).max() <= SR_REGRESSION_GATE
        model.debug_set(dev, "skip_dead_views", 0)
        dense = model(tl, ta).cpu().numpy()
        assert np.array_equal(sr, dense)


def test_dead_view_lists_large_batch(net, dev):
    """B * L large enough that the liveness scratch leaves shared memory (global-memory path of live_lists_kernel):
    every imageset of the batch must equal the same imageset run alone."""
    rng = np.random.RandomState(11)
    b, l = 300, 32
    lrs = torch.from_numpy(rng.rand(b, l, 16, 16).astype(np.float32)).to(dev)
    real = rng.randint(1, l + 1, size=b)
    alphas = torch.zeros(b, l, device=dev)
    for i, n in enumerate(real):
        alphas[i, :n] = 1
        lrs[i, n:] = 0
    sr = net(lrs, alphas)
    for i in (0, 1, 57, 150, 299):
        assert torch.equal(sr[i:i + 1], net(lrs[i:i + 1], alphas[i:i + 1])), i


def test_dead_view_skipping_saves_time(hb, net, dev):
    """B8 L32 64x64 with 12 real views per imageset (config.json n_views = 32, Proba-V scenes average 19 views): the
    padded run must be clearly cheaper than the all-real run."""
    lrs = torch.rand(8, 32, 64, 64, device=dev)
    full = torch.ones(8, 32, device=dev)
    padded = full.clone()
    padded[:, 12:] = 0
    lrs_p = lrs.clone()
    lrs_p[:, 12:] = 0

    def timed(x, a):
        for _ in range(3):
            net(x, a)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            net(x, a)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / 10

    t_full, t_pad = timed(lrs, full), timed(lrs_p, padded)
    assert t_pad < 0.7 * t_full, (t_full, t_pad)


def test_full_size_properties_c2(net, dev):
    """BASELINE.json configs[1] size (B32 L16 128x128): size-independent properties.
    Imagesets are independent => each batch row equals the same imageset run alone (bit-exact:
    the kernels are deterministic), a batch permutation permutes the output, and one row is
    checked against the oracle."""
    g = torch.Generator().manual_seed(5)
    lrs = torch.rand(32, 16, 128, 128, generator=g)
    alphas = torch.ones(32, 16)
    alphas[3, 11:] = 0
    lrs[3, 11:] = 0
    tl, ta = lrs.to(dev), alphas.to(dev)
    sr = net(tl, ta)
    assert tuple(sr.shape) == (32, 1, 384, 384) and torch.isfinite(sr).all()
    for i in (0, 3, 31):
        alone = net(tl[i:i + 1], ta[i:i + 1])
        assert torch.equal(alone[0], sr[i])
    perm = torch.randperm(32, generator=g)
    sr_p = net(tl[perm.to(dev)], ta[perm.to(dev)])
    assert torch.equal(sr_p, sr[perm.to(dev)])
    again = net(tl, ta)
    assert torch.equal(again, sr)                                        # run-to-run determinism
    params = hrnet_oracle.make_params(cases.WEIGHT_SEED)
    ref = hrnet_oracle.hrnet_forward(params, lrs[3:4].numpy(), alphas[3:4].numpy()).numpy()
    err = np.abs(sr[3:4].cpu().numpy() - ref).max()
    assert err <= SR_GATE and err <= SR_REGRESSION_GATE, err


def test_row_partition_does_not_change_results(hb, dev):
    """The conv kernels split the flattened (image, row) space evenly over the CTAs; odd CTA counts put strip
    boundaries in the middle of images (1-row strips included).  Results must be bit-identical."""
    lrs, alphas = cases.hrnet_inputs("b2_l9_s16")
    tl, ta = torch.from_numpy(lrs).to(dev), torch.from_numpy(alphas).to(dev)
    outs = []
    for ctas in (0, 1, 2, 7, 37, 146):
        for wave in (0, 1):                     # encoder convs always; the fusion convs in the three-launch schedule
            model = hb.HRNet(hrnet_oracle.DEFAULT_NETWORK_CONFIG).eval()
            model.load_state_dict(hrnet_oracle.make_params(cases.WEIGHT_SEED))
            model = model.to(dev)
            model.debug_set(dev, "max_ctas", ctas)
            model.debug_set(dev, "fuse_wave", wave)
            outs.append(model(tl, ta))
    for split in (3, 16, 1000):             # several short ranges per CTA, dealt round-robin (1000: mostly empty ranges)
        model.debug_set(dev, "max_ctas", 5)
        model.debug_set(dev, "strip_split", split)
        outs.append(model(tl, ta))
    for o in outs[1:]:
        assert torch.equal(o, outs[0])


@pytest.mark.parametrize("b,l,s", [(1, 1, 8), (1, 2, 16), (2, 3, 33), (2, 4, 128), (1, 5, 100), (3, 2, 1), (2, 2, 2), (1, 16, 64)])
def test_fused_resblock_is_bit_identical_to_two_launches(hb, dev, b, l, s):
    """resblock64_umma (both convs of an encoder ResidualBlock in one launch, the intermediate rows never leave the SM)
    rounds the intermediate to bf16 exactly like the stand-alone layer: the SR image must equal the two-launch path bit
    for bit, for any CTA count (strip boundaries move, two extra intermediate rows per strip are recomputed), with
    alpha = 0 views skipped or not."""
    model = hb.HRNet(hrnet_oracle.DEFAULT_NETWORK_CONFIG).eval()
    model.load_state_dict(hrnet_oracle.make_params(cases.WEIGHT_SEED))
    model = model.to(dev)
    g = torch.Generator().manual_seed(1000 + 10 * s + l)
    lrs = torch.rand(b, l, s, s, generator=g).to(dev)
    al = torch.ones(b, l)
    if l > 2:
        al[0, -1] = 0.0
    al = al.to(dev)
    outs = []
    for fuse in (1, 0):
        for ctas in (0, 3, 37):
            model.debug_set(dev, "fuse_resblock", fuse)
            model.debug_set(dev, "max_ctas", ctas)
            outs.append(model(lrs, al).clone())
    for o in outs[1:]:
        assert torch.equal(o, outs[0])
    ref = hrnet_oracle.hrnet_forward(hrnet_oracle.make_params(cases.WEIGHT_SEED), lrs.cpu().numpy(), al.cpu().numpy()).numpy()
    assert np.abs(outs[0].cpu().numpy() - ref).max() <= SR_GATE


def _net_with(hb, dev, **knobs):
    model = hb.HRNet(hrnet_oracle.DEFAULT_NETWORK_CONFIG).eval()
    model.load_state_dict(hrnet_oracle.make_params(cases.WEIGHT_SEED))
    model = model.to(dev)
    for knob, value in knobs.items():
        model.debug_set(dev, knob, value)
    return model


@pytest.mark.parametrize("b,l,s", [(1, 2, 8), (1, 2, 16), (2, 3, 33), (2, 4, 128), (1, 5, 100), (3, 2, 1), (2, 2, 2), (1, 16, 64),
                                   (4, 8, 128), (3, 9, 40), (2, 32, 32)])
def test_wavefront_fusion_is_bit_identical_to_three_launches(hb, dev, b, l, s):
    """fuse_wave_umma (one launch per fusion level: conv A / B / C as five-CTA streams that hand rows over through
    L2-resident rings, HRNet.py:93-97, 113-128) against the three-launch schedule: every output element sees the same
    products in the same order and the same rounding points, so the SR image must be bit-identical -- with full view
    sets, padded / holed alpha patterns (carried views), and whatever the stream partition and ring depth."""
    g = torch.Generator().manual_seed(b * 1000 + l * 10 + s)
    lrs = torch.rand(b, l, s, s, generator=g).to(dev)
    alphas = torch.ones(b, l, device=dev)
    ref_net = _net_with(hb, dev, fuse_wave=0)
    before = hb.kernel_launch_count()
    ref = ref_net(lrs, alphas)
    launches_ref = hb.kernel_launch_count() - before
    wave_net = _net_with(hb, dev, fuse_wave=1)
    before = hb.kernel_launch_count()
    out = wave_net(lrs, alphas)
    launches_wave = hb.kernel_launch_count() - before
    levels = int(np.floor(np.log2(l)))
    assert launches_ref - launches_wave == 2 * levels            # one launch per level instead of three
    assert torch.equal(out, ref)
    patterns = [alphas.clone() for _ in range(3)]
    if l > 2:
        patterns[0][:, l - 1:] = 0                                # trailing padding (utils.py:92-95)
        patterns[1][0, 1] = 0                                     # a hole: some pair loses its bob at level 0
        patterns[1][-1, l // 2:] = 0
        patterns[2][:, 1:] = 0                                    # only view 0 left: every level just carries alice
        for al in patterns:
            assert torch.equal(wave_net(lrs * al[:, :, None, None], al), ref_net(lrs * al[:, :, None, None], al))
    for knobs in ({"wave_streams": 1}, {"wave_streams": 3, "wave_ring_rows": 8}, {"wave_streams": 7, "wave_publish_rows": 3, "wave_ring_rows": 12},
                  {"wave_ring_rows": 64}):
        other = _net_with(hb, dev, fuse_wave=1, **knobs)
        assert torch.equal(other(lrs, alphas), ref), knobs
        if l > 2:
            assert torch.equal(other(lrs * patterns[1][:, :, None, None], patterns[1]), ref_net(lrs * patterns[1][:, :, None, None], patterns[1])), knobs


@pytest.mark.parametrize("b,l,s", [(1, 2, 8), (2, 3, 33), (2, 4, 128), (1, 5, 100), (3, 2, 1), (1, 16, 64), (4, 8, 128)])
def test_encoder_wavefront_is_bit_identical_to_per_layer_launches(hb, dev, b, l, s):
    """enc_wave_umma (opt-in, knob "enc_wave"): the two ResidualBlocks and the final conv of the encoder (HRNet.py:55-60) as
    five-CTA streams over L2 rings, with x1 read twice (conv input of ResidualBlock 1 and, rows later, its skip connection).
    Same products, same order, same rounding points as resblock64_umma / conv3x3_umma<64>: bit-identical SR, with dead views
    and whatever the stream partition and ring depth."""
    g = torch.Generator().manual_seed(7 * b + 100 * l + s)
    lrs = torch.rand(b, l, s, s, generator=g).to(dev)
    alphas = torch.ones(b, l, device=dev)
    if l > 2:
        alphas[-1, l - 1:] = 0
    ref = _net_with(hb, dev, fuse_wave=0, enc_wave=0)(lrs, alphas)
    for knobs in ({"enc_wave": 1, "fuse_wave": 0}, {"enc_wave": 1}, {"enc_wave": 1, "wave_streams": 1, "enc_ring_rows": 12},
                  {"enc_wave": 1, "wave_streams": 7, "enc_ring_rows": 40}):
        model = _net_with(hb, dev, **knobs)
        before = hb.kernel_launch_count()
        out = model(lrs, alphas)
        assert torch.equal(out, ref), knobs
        assert torch.equal(model(lrs, alphas), ref), knobs          # second forward: counters were reset
    assert hb.kernel_launch_count() > before


def test_wavefront_fusion_soak_c2(hb, dev):
    """Hand-over protocol under load: 120 back-to-back forwards at BASELINE configs[1] size on three alternating inputs (the
    default ring depth and a shallow one) must reproduce the three-launch results every time: a missed release, a stale ring
    row or an unordered proxy shows up as a changed bit (a single-buffered residual slot without its proxy fence failed
    this check in one forward out of ten)."""
    g = torch.Generator().manual_seed(99)
    xs = [torch.rand(32, 16, 128, 128, generator=g).to(dev) for _ in range(3)]
    alphas = torch.ones(32, 16, device=dev)
    ref_net = _net_with(hb, dev, fuse_wave=0)
    refs = [ref_net(x, alphas) for x in xs]
    for knobs in ({}, {"wave_ring_rows": 10}):
        wave_net = _net_with(hb, dev, fuse_wave=1, **knobs)
        bad = [i for i in range(60) if not torch.equal(wave_net(xs[i % 3], alphas), refs[i % 3])]
        assert not bad, (knobs, bad)


def test_wavefront_fusion_two_handles_two_streams(hb, dev):
    """Wavefront launches spin on flags of sibling CTAs, so two of them must never share the GPU half and half: the
    library serialises them per device.  Two models on two streams, interleaved, must both finish and agree."""
    g = torch.Generator().manual_seed(5)
    x1, x2 = torch.rand(8, 8, 128, 128, generator=g).to(dev), torch.rand(8, 8, 128, 128, generator=g).to(dev)
    alphas = torch.ones(8, 8, device=dev)
    m1, m2 = _net_with(hb, dev), _net_with(hb, dev)
    r1, r2 = m1(x1, alphas), m2(x2, alphas)
    torch.cuda.synchronize()
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    outs = []
    for _ in range(10):
        with torch.cuda.stream(s1):
            outs.append((m1(x1, alphas), r1))
        with torch.cuda.stream(s2):
            outs.append((m2(x2, alphas), r2))
    torch.cuda.synchronize()
    assert all(torch.equal(o, r) for o, r in outs)


@pytest.mark.parametrize("b,l,s", [(1, 2, 8), (2, 3, 33), (2, 4, 128), (1, 5, 100), (3, 2, 1), (1, 16, 64), (2, 8, 256)])
def test_multicast_cluster_pairs_are_bit_identical(hb, dev, b, l, s):
    """The 128 -> 128 convs of the fusion stage run as clusters of two CTAs (the two 64-channel output halves of the same
    rows) that each fetch one K chunk of every input row and multicast it to both; the plain launch ("mcast" = 0) must
    give the same image bit for bit, for any CTA count and with dead views skipped through the live-work lists."""
    model = hb.HRNet(hrnet_oracle.DEFAULT_NETWORK_CONFIG).eval()
    model.load_state_dict(hrnet_oracle.make_params(cases.WEIGHT_SEED))
    model = model.to(dev)
    model.debug_set(dev, "fuse_wave", 0)        # the three-launch schedule is what uses these convolutions when W <= 128
    g = torch.Generator().manual_seed(2000 + 10 * s + l)
    lrs = torch.rand(b, l, s, s, generator=g).to(dev)
    al = (torch.rand(b, l, generator=g) > 0.25).float()
    al[:, 0] = 1.0
    al = al.to(dev)
    outs = []
    for mc in (1, 0):
        for ctas in (0, 6, 37):
            model.debug_set(dev, "mcast", mc)
            model.debug_set(dev, "max_ctas", ctas)
            outs.append(model(lrs, al).clone())
    for o in outs[1:]:
        assert torch.equal(o, outs[0])
    model.debug_set(dev, "max_ctas", 0)
    model.debug_set(dev, "mcast", 2)            # "required": errors out if the cluster path were silently skipped on this GPU
    assert torch.equal(model(lrs, al), outs[0])


def test_forward_host_equals_device_path(net, dev):
    lrs, alphas = cases.hrnet_inputs("b2_l4_s32")
    a = net(torch.from_numpy(lrs).to(dev), torch.from_numpy(alphas).to(dev)).cpu()
    b = net.forward_host(torch.from_numpy(lrs).pin_memory(), torch.from_numpy(alphas).pin_memory(), device=dev)
    assert torch.equal(a, b)


def test_forward_host_pipeline_depths(hb, dev):
    """hrn_forward_host cuts the batch into chunks whose copies overlap compute; any depth gives identical output."""
    model = hb.HRNet(hrnet_oracle.DEFAULT_NETWORK_CONFIG).eval()
    model.load_state_dict(hrnet_oracle.make_params(cases.WEIGHT_SEED))
    model = model.to(dev)
    g = torch.Generator().manual_seed(8)
    lrs, alphas = torch.rand(7, 4, 32, 32, generator=g).pin_memory(), torch.ones(7, 4).pin_memory()
    ref = model(lrs.to(dev), alphas.to(dev)).cpu()
    for depth in (1, 2, 3, 8):
        model.debug_set(dev, "host_chunks", depth)
        assert torch.equal(model.forward_host(lrs, alphas, device=dev), ref), depth
    model.debug_set(dev, "host_chunks", 0)
    assert torch.equal(model.forward_host(lrs.clone(), alphas.clone(), device=dev), ref)    # pageable host memory too


def test_forward_host_submit_wait_pipeline(hb, dev):
    """hrn_forward_host_submit / _wait: batches of different shapes in flight two at a time, waited in and out of
    order, give exactly what the device path gives; a retired ticket can still be waited; unknown tickets raise."""
    model = hb.HRNet(hrnet_oracle.DEFAULT_NETWORK_CONFIG).eval()
    model.load_state_dict(hrnet_oracle.make_params(cases.WEIGHT_SEED))
    model = model.to(dev)
    g = torch.Generator().manual_seed(31)
    shapes = [(3, 4, 32), (2, 5, 40), (1, 16, 24), (4, 2, 32), (2, 3, 33), (3, 4, 32), (1, 1, 16)]
    batches = []
    for (b, l, s) in shapes:
        al = (torch.rand(b, l, generator=g) > 0.2).float()
        al[:, 0] = 1.0
        batches.append((torch.rand(b, l, s, s, generator=g).pin_memory(), al.pin_memory()))
    refs = [model(x.to(dev), a.to(dev)).cpu() for x, a in batches]
    outs = list(model.forward_host_iter(batches, device=dev))
    assert len(outs) == len(refs)
    for o, r in zip(outs, refs):
        assert torch.equal(o, r)
    # four submits without a wait: the third and fourth retire the first and second
    pend = [model.forward_host_submit(*batches[i], device=dev) for i in range(4)]
    for i in (3, 0, 2, 1):
        assert torch.equal(model.forward_host_wait(pend[i]), refs[i]), i
    assert torch.equal(model.forward_host_wait(pend[0]), refs[0])          # waiting twice is harmless
    # mixes with the synchronous call on the same stream
    p = model.forward_host_submit(*batches[4], device=dev)
    assert torch.equal(model.forward_host(*batches[5], device=dev), refs[5])
    assert torch.equal(model.forward_host_wait(p), refs[4])
    with pytest.raises(RuntimeError):
        model.forward_host_wait((p[0], 10 ** 6, p[2]))
    # pageable host memory works too (no overlap then)
    q = model.forward_host_submit(batches[6][0].clone(), batches[6][1].clone(), out_host=torch.empty(1, 1, 48, 48), device=dev)
    assert torch.equal(model.forward_host_wait(q), refs[6])


def test_workspace_cap_slices_the_batch(hb, dev):
    """A batch whose activation workspace exceeds the cap is run in slices; the output must not change."""
    model = hb.HRNet(hrnet_oracle.DEFAULT_NETWORK_CONFIG).eval()
    model.load_state_dict(hrnet_oracle.make_params(cases.WEIGHT_SEED))
    model = model.to(dev)
    g = torch.Generator().manual_seed(21)
    lrs, alphas = torch.rand(7, 4, 48, 48, generator=g).to(dev), torch.ones(7, 4, device=dev)
    ref = model(lrs, alphas)
    model.debug_set(dev, "workspace_mb", 12)      # one imageset needs 5 * 4 * 48 * 48 * 128 B = 5.6 MB -> slices of 2
    assert torch.equal(model(lrs, alphas), ref)
    model.debug_set(dev, "workspace_mb", 1)       # smaller than one imageset: falls back to one imageset per slice
    assert torch.equal(model(lrs, alphas), ref)


def test_forward_on_side_stream(net, dev):
    """All work is enqueued on the caller's current stream (torch.cuda.current_stream)."""
    lrs, alphas = cases.hrnet_inputs("b2_l4_s32")
    tl, ta = torch.from_numpy(lrs).to(dev), torch.from_numpy(alphas).to(dev)
    ref = net(tl, ta)
    torch.cuda.synchronize()
    side = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(side):
        big = torch.rand(4096, 4096, device=dev) @ torch.rand(4096, 4096, device=dev)      # keep the side stream busy first
        out = net(tl, ta)
        done = torch.cuda.Event()
        done.record(side)
    done.synchronize()
    assert torch.equal(out, ref) and big.shape == (4096, 4096)


def test_two_handles_on_two_streams(hb, dev):
    """Handles own their workspace and live-work lists: two models with different weights, interleaved on two streams,
    must each reproduce their own sequential result bit for bit."""
    lrs, alphas = cases.hrnet_inputs("b2_l9_s16")
    tl, ta = torch.from_numpy(lrs).to(dev), torch.from_numpy(alphas).to(dev)
    models = []
    for seed in (0, 1):
        m = hb.HRNet(hrnet_oracle.DEFAULT_NETWORK_CONFIG).eval()
        m.load_state_dict(hrnet_oracle.make_params(seed))
        models.append(m.to(dev))
    refs = [m(tl, ta).clone() for m in models]
    assert not torch.equal(refs[0], refs[1])
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream(device=dev) for _ in models]
    outs = [[], []]
    for _ in range(8):
        for k, (m, st) in enumerate(zip(models, streams)):
            with torch.cuda.stream(st):
                outs[k].append(m(tl, ta))
    torch.cuda.synchronize()
    for k in range(2):
        for o in outs[k]:
            assert torch.equal(o, refs[k])


def test_one_handle_two_streams_are_ordered(hb, dev):
    """two forwards of ONE module on different torch streams share the handle's workspace; the handle orders
    them with an event, so both results must equal the single-stream results."""
    model = hb.HRNet(hrnet_oracle.DEFAULT_NETWORK_CONFIG).eval()
    model.load_state_dict(hrnet_oracle.make_params(cases.WEIGHT_SEED))
    model = model.to(dev)
    g = torch.Generator().manual_seed(77)
    a = (torch.rand(4, 8, 96, 96, generator=g).to(dev), torch.ones(4, 8, device=dev))
    b = (torch.rand(4, 8, 96, 96, generator=g).to(dev), torch.ones(4, 8, device=dev))
    ref_a, ref_b = model(*a), model(*b)
    torch.cuda.synchronize()
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    for _ in range(5):
        with torch.cuda.stream(s1):
            out_a = model(*a)
        with torch.cuda.stream(s2):
            out_b = model(*b)
        torch.cuda.synchronize()
        assert torch.equal(out_a, ref_a) and torch.equal(out_b, ref_b)


def test_reserve_presizes_the_workspace(hb, dev):
    model = hb.HRNet(hrnet_oracle.DEFAULT_NETWORK_CONFIG).eval()
    model.load_state_dict(hrnet_oracle.make_params(cases.WEIGHT_SEED))
    model = model.to(dev)
    model.reserve(dev, 4, 8, 64, 64)
    free0, _ = torch.cuda.mem_get_info(dev)
    lrs, alphas = torch.rand(4, 8, 64, 64, device=dev), torch.ones(4, 8, device=dev)
    sr = model(lrs, alphas)
    torch.cuda.synchronize()
    free1, _ = torch.cuda.mem_get_info(dev)
    assert free0 - free1 < 5 * 4 * 8 * 64 * 64 * 128 // 2          # the forward did not allocate the activation buffers again
    assert torch.isfinite(sr).all()
    with pytest.raises(RuntimeError):
        model.reserve(dev, 0, 8, 64, 64)


def test_data_edits_need_invalidate_or_verify(hb, dev):
    """in-place edits through .data do not bump torch's version counter; invalidate_weights() (or
    verify_weights) makes the next forward upload the new values."""
    model = hb.HRNet(hrnet_oracle.DEFAULT_NETWORK_CONFIG).eval()
    model.load_state_dict(hrnet_oracle.make_params(cases.WEIGHT_SEED))
    model = model.to(dev)
    lrs, alphas = torch.rand(1, 4, 32, 32, device=dev), torch.ones(1, 4, device=dev)
    base = model(lrs, alphas).clone()
    model.decode.final.bias.data.add_(0.25)
    model.invalidate_weights()
    moved = model(lrs, alphas)
    assert torch.allclose(moved, base + 0.25, atol=1e-6)
    model.verify_weights = True
    model(lrs, alphas)
    model.decode.final.bias.data.sub_(0.25)
    assert torch.allclose(model(lrs, alphas), base, atol=1e-6)
    import copy
    twin = copy.deepcopy(model)                                    # after a forward: the handles stay behind
    assert torch.equal(twin(lrs, alphas), model(lrs, alphas))


def test_forward_rejects_bad_inputs(hb, net, dev):
    with pytest.raises(RuntimeError):
        net(torch.rand(1, 2, 16, 16), torch.ones(1, 2))                  # CPU tensors: no fallback
    with pytest.raises(ValueError):
        net(torch.rand(1, 2, 16, 24, device=dev), torch.ones(1, 2, device=dev))   # non-square (HRNet.py:204)
    fresh = hb.HRNet(hrnet_oracle.DEFAULT_NETWORK_CONFIG).to(dev)
    with pytest.raises(RuntimeError):
        fresh(torch.rand(1, 2, 16, 16, device=dev), torch.ones(1, 2, device=dev))  # train mode + grad: unsupported
    with torch.no_grad():
        out = fresh(torch.rand(1, 2, 16, 16, device=dev), torch.ones(1, 2, device=dev))
    assert tuple(out.shape) == (1, 1, 48, 48)


def test_state_dict_roundtrip_changes_output(hb, dev):
    """load_state_dict (predict.py:98-99) must reach the device copy of the weights."""
    model = hb.HRNet(hrnet_oracle.DEFAULT_NETWORK_CONFIG).eval().to(dev)
    x, a = torch.rand(1, 2, 16, 16, device=dev), torch.ones(1, 2, device=dev)
    y0 = model(x, a)
    model.load_state_dict(hrnet_oracle.make_params(123))
    y1 = model(x, a)
    ref = hrnet_oracle.hrnet_forward(hrnet_oracle.make_params(123), x.cpu().numpy(), a.cpu().numpy())
    assert not torch.equal(y0, y1)
    assert (y1.cpu() - ref).abs().max() <= SR_REGRESSION_GATE


# ---------------------------------------------------------------------------- Lanczos
def test_lanczos_taps_known_answers(hb, dev, golden):
    d = torch.tensor(cases.LANCZOS_TAP_SHIFTS, dtype=torch.float32, device=dev).view(-1, 1)
    taps = hb.lanczos_kernel(d, a=3, N=7).cpu().numpy()
    assert np.abs(taps - golden["lanczos"]["taps"]).max() <= 5e-7
    assert np.allclose(taps.sum(1), 1.0, atol=1e-6)


@pytest.mark.parametrize("name", list(cases.LANCZOS_CASES))
def test_lanczos_shift_matches_reference_golden(hb, dev, golden, name):
    img, shift, p = cases.lanczos_inputs(name)
    out = hb.lanczos_shift(torch.from_numpy(img).to(dev), torch.from_numpy(shift).to(dev), p=p, a=3, N=7)
    assert tuple(out.shape) == img.shape
    out = out.cpu().numpy()
    assert np.abs(out - golden["lanczos"][name]).max() <= LANCZOS_GATE
    assert np.abs(out - scoring_oracle.lanczos_shift(img, shift, p=p)).max() <= LANCZOS_GATE


@pytest.mark.parametrize("shape", [(1, 3, 8, 8, 3), (2, 2, 37, 96, 5), (1, 2, 70, 132, 5), (1, 3, 33, 260, 4), (1, 2, 130, 128, 1),
                                   (1, 2, 64, 12, 2), (3, 2, 5, 384, 4), (1, 2, 150, 37, 5), (1, 40, 384, 384, 5)])
def test_lanczos_both_kernels_match_oracle(hb, dev, shape):
    """The N = 7 path has two kernels (TMA-fed tiles for 16-byte aligned rows, a register-window kernel otherwise): both
    against the oracle on shapes that put the reflect ring, the zero ring (p < 3) and the image edge inside, at the edge of
    and across tiles -- and against each other bit for bit where both apply (same fmaf order)."""
    nb, c, h, w, p = shape
    rng = np.random.RandomState(h * 1000 + w)
    img = rng.rand(nb, c, h, w).astype(np.float32)
    shift = rng.uniform(-1, 1, size=(c, 2)).astype(np.float32)
    ref = scoring_oracle.lanczos_shift(img, shift, p=p)
    outs = []
    try:
        for knob in (0, 1):
            hb.scoring_debug_set("lanczos_scalar", knob)
            out = hb.lanczos_shift(torch.from_numpy(img).to(dev), torch.from_numpy(shift).to(dev), p=p, a=3, N=7)
            outs.append(out.cpu().numpy())
            assert np.abs(outs[-1] - ref).max() <= LANCZOS_GATE, f"knob {knob}"
    finally:
        hb.scoring_debug_set("lanczos_scalar", 0)
    assert np.array_equal(outs[0], outs[1])


def test_lanczos_unaligned_view_falls_back(hb, dev):
    """A storage offset of one float makes the rows 4-byte aligned only: the launcher must take the other kernel, not fault."""
    rng = np.random.RandomState(5)
    flat = torch.from_numpy(rng.rand(1 + 2 * 64 * 64).astype(np.float32)).to(dev)
    img = flat[1:].view(1, 2, 64, 64)
    shift = torch.tensor([[0.25, -0.5], [-0.75, 0.1]], device=dev)
    out = hb.lanczos_shift(img, shift, p=5)
    ref = scoring_oracle.lanczos_shift(img.cpu().numpy(), shift.cpu().numpy(), p=5)
    assert np.abs(out.cpu().numpy() - ref).max() <= LANCZOS_GATE


def test_lanczos_properties_full_size(hb, dev):
    """384x384 SR size: linearity, and an integer shift is a plain translation away from the border."""
    g = torch.Generator().manual_seed(9)
    x = torch.rand(1, 32, 384, 384, generator=g).to(dev)
    y = torch.rand(1, 32, 384, 384, generator=g).to(dev)
    shift = (torch.rand(32, 2, generator=g) * 2 - 1).to(dev)
    lin = hb.lanczos_shift(2.0 * x - 0.5 * y, shift, p=5)
    sep = 2.0 * hb.lanczos_shift(x, shift, p=5) - 0.5 * hb.lanczos_shift(y, shift, p=5)
    assert (lin - sep).abs().max() <= 1e-5
    ones = torch.ones(32, 2, device=dev)
    moved = hb.lanczos_shift(x, ones, p=5)                    # d = +1 samples the image at y+1, x+1
    assert (moved[..., 4:-4, 4:-4] - x[..., 5:-3, 5:-3]).abs().max() <= 1e-5


def test_lanczos_rejects_bad_arguments(hb, dev):
    img = torch.rand(1, 2, 8, 8, device=dev)
    with pytest.raises(RuntimeError):
        hb.lanczos_shift(img, torch.zeros(2, 2, device=dev), p=8)        # ReflectionPad2d needs p < size
    with pytest.raises(RuntimeError):
        hb.lanczos_shift(img, torch.zeros(2, 2, device=dev), N=6)        # even kernel width
    with pytest.raises(RuntimeError):
        hb.lanczos_shift(img.cpu(), torch.zeros(2, 2))


# ---------------------------------------------------------------------------- cPSNR shift search
@pytest.mark.parametrize("name", list(cases.CPSNR_CASES))
def test_shift_cpsnr_matches_reference_golden(hb, golden, name):
    sr, hr, hm = cases.cpsnr_inputs(name)
    g = golden["cpsnr"]
    best, xy, table = hb.shift_cPSNR_argmax(sr, hr, hm, border_w=3)
    assert best.dtype == np.float32 and best.shape == (sr.shape[0],)
    ref_max, ref_arg, ref_sites = g[name + "__max"], g[name + "__argmax"], g[name + "__sites"]
    assert np.array_equal(xy[:, 0] * 7 + xy[:, 1], ref_arg)                       # best shift: bit-exact
    assert np.array_equal(np.isnan(table), np.isnan(ref_sites))
    assert np.array_equal(np.isposinf(table), np.isposinf(ref_sites))
    fin = np.isfinite(ref_sites)
    assert np.abs(table[fin] - ref_sites[fin]).max(initial=0.0) <= CPSNR_KERNEL_GATE_DB
    finm = np.isfinite(ref_max)
    assert np.abs(best[finm] - ref_max[finm]).max(initial=0.0) <= CPSNR_KERNEL_GATE_DB
    assert np.array_equal(np.isnan(best), np.isnan(ref_max)) and np.array_equal(np.isposinf(best), np.isposinf(ref_max))
    # 2-D call signature of the reference (Evaluator.py:21-24): scalar out
    with np.errstate(all="ignore"):
        one = hb.shift_cPSNR(sr[0], hr[0], hm[0])
    assert np.ndim(one) == 0
    if np.isfinite(ref_max[0]):
        assert abs(one - ref_max[0]) <= CPSNR_KERNEL_GATE_DB


def test_cpsnr_plain_and_uint16(hb):
    rng = np.random.RandomState(21)
    sr, hr = rng.rand(2, 50, 50).astype(np.float32), rng.rand(2, 50, 50).astype(np.float32)
    hm = (rng.rand(2, 50, 50) > 0.2).astype(np.float32)
    assert np.abs(hb.cPSNR(sr, hr, hm) - scoring_oracle.cpsnr(sr, hr, hm)).max() <= CPSNR_KERNEL_GATE_DB
    sr16, hr16 = (sr * 65535).astype(np.uint16), (hr * 65535).astype(np.uint16)
    assert np.abs(hb.cPSNR(sr16, hr16, hm) - scoring_oracle.cpsnr(sr16, hr16, hm)).max() <= CPSNR_KERNEL_GATE_DB
    with pytest.raises(AssertionError):
        hb.cPSNR(sr + 1.0, hr, hm)                                              # Evaluator.py:30


def test_shift_cpsnr_known_shift_full_size(hb, dev):
    """384x384, batch 32 on device: hr = roll(sr, (ry, rx)) -> best site (3+ry, 3+rx), cMSE = 0 -> +inf."""
    g = torch.Generator().manual_seed(4)
    sr = torch.rand(32, 384, 384, generator=g)
    shifts = torch.randint(-3, 4, (32, 2), generator=g)
    hr = torch.stack([torch.roll(sr[i], (int(shifts[i, 0]), int(shifts[i, 1])), (0, 1)) for i in range(32)])
    hm = (torch.rand(32, 384, 384, generator=g) > 0.1).float()
    best, xy, _ = hb.shift_cPSNR_argmax(sr.to(dev), hr.to(dev), hm.to(dev))
    assert best.is_cuda and torch.isinf(best).all()
    assert torch.equal(xy.cpu().long(), shifts + 3)


@pytest.mark.parametrize("b,s,kind", [(2, 8, "rand"), (3, 12, "rand"), (2, 20, "soft"), (5, 136, "rand"), (2, 260, "shift"),
                                      (3, 384, "rand"), (40, 64, "rand"), (3, 40, "degenerate")])
def test_shift_cpsnr_window_kernel_vs_generic_and_oracle(hb, dev, b, s, kind):
    """border_w = 3 runs the 49-site window kernel (one warp sweeps a band of rows for all sites); the general
    shift-window kernel behind the "cpsnr_generic" knob and the oracle must give the same table and the same best shift.
    Sizes cover one lane of columns, partly filled column blocks, several bands and the degenerate masks."""
    rng = np.random.RandomState(100 + s + b)
    sr = rng.rand(b, s, s).astype(np.float32)
    hr = rng.rand(b, s, s).astype(np.float32)
    hm = (rng.rand(b, s, s) > 0.1).astype(np.float32)
    if kind == "soft":
        hm = rng.rand(b, s, s).astype(np.float32)
    if kind == "shift":
        hr = np.clip(np.roll(sr, (2, -1), (1, 2)) + 0.02 + 0.01 * rng.randn(b, s, s), 0, 1).astype(np.float32)
    if kind == "degenerate":
        hm[0] = 0.0
        hr[1] = sr[1]
        hm[2] = 0.0
        hm[2, s // 2, s // 2] = 1.0
    args = [torch.from_numpy(a).to(dev) for a in (sr, hr, hm)]
    best_w, xy_w, tab_w = hb.shift_cPSNR_argmax(*args)                # default: the one-pass kernel (+ fallback for flagged sites)
    hb.scoring_debug_set("cpsnr_onepass", 0)                         # the two-pass window kernels, variant by batch size
    try:
        best_t, xy_t, tab_t = hb.shift_cPSNR_argmax(*args)
    finally:
        hb.scoring_debug_set("cpsnr_onepass", 1)
    assert torch.equal(xy_t, xy_w)
    assert torch.equal(torch.isnan(tab_t), torch.isnan(tab_w)) and torch.equal(torch.isinf(tab_t), torch.isinf(tab_w))
    assert np.abs(np.nan_to_num(tab_t.cpu().numpy() - tab_w.cpu().numpy(), nan=0.0, posinf=0.0, neginf=0.0)).max() <= CPSNR_KERNEL_GATE_DB
    hb.scoring_debug_set("cpsnr_generic", 1)
    try:
        best_g, xy_g, tab_g = hb.shift_cPSNR_argmax(*args)
    finally:
        hb.scoring_debug_set("cpsnr_generic", 0)
    # the packed fp32x2 window kernel (x split over two warps; the default is the scalar 49-sites-per-warp kernel) and
    # the chunked pass-1 / pass-2 schedule
    hb.scoring_debug_set("cpsnr_window_v1", 0)
    try:
        best_1, xy_1, tab_1 = hb.shift_cPSNR_argmax(*args)
        hb.scoring_debug_set("cpsnr_window_v1", 2)                # x split over two warps, scalar fp32
        best_2, xy_2, tab_2 = hb.shift_cPSNR_argmax(*args)
        assert torch.equal(xy_2, xy_1)
        assert np.abs(np.nan_to_num(tab_2.cpu().numpy() - tab_1.cpu().numpy(), nan=0.0, posinf=0.0, neginf=0.0)).max() <= CPSNR_KERNEL_GATE_DB
        hb.scoring_debug_set("cpsnr_window_v1", 1)                # the 49-sites-per-warp kernel, whatever the batch size
        best_3, xy_3, tab_3 = hb.shift_cPSNR_argmax(*args)
        assert torch.equal(xy_3, xy_1)
        assert np.abs(np.nan_to_num(tab_3.cpu().numpy() - tab_1.cpu().numpy(), nan=0.0, posinf=0.0, neginf=0.0)).max() <= CPSNR_KERNEL_GATE_DB
    finally:
        hb.scoring_debug_set("cpsnr_window_v1", -1)
    hb.scoring_debug_set("cpsnr_chunk", 2)
    try:
        best_c, xy_c, tab_c = hb.shift_cPSNR_argmax(*args)
    finally:
        hb.scoring_debug_set("cpsnr_chunk", 0)
    assert torch.equal(xy_c, xy_w) and torch.equal(xy_1, xy_w)
    assert np.abs(np.nan_to_num(tab_c.cpu().numpy() - tab_w.cpu().numpy(), nan=0.0, posinf=0.0, neginf=0.0)).max() <= CPSNR_KERNEL_GATE_DB
    ref_max, ref_arg, ref_sites = scoring_oracle.shift_cpsnr(sr, hr, hm)
    tab_w, tab_g, tab_1, ref_sites = tab_w.cpu().numpy(), tab_g.cpu().numpy(), tab_1.cpu().numpy(), ref_sites.T
    for tab in (tab_w, tab_g, tab_1):
        assert np.array_equal(np.isnan(tab), np.isnan(ref_sites))
        assert np.array_equal(np.isposinf(tab), np.isposinf(ref_sites))
        fin = np.isfinite(ref_sites)
        assert np.abs(tab[fin] - ref_sites[fin]).max(initial=0.0) <= CPSNR_KERNEL_GATE_DB
    arg_w = (xy_w[:, 0] * 7 + xy_w[:, 1]).cpu().numpy()
    arg_g = (xy_g[:, 0] * 7 + xy_g[:, 1]).cpu().numpy()
    assert np.array_equal(arg_w, ref_arg) and np.array_equal(arg_g, ref_arg)
    assert torch.equal(torch.isnan(best_w), torch.isnan(best_g))
    # run to run: fixed-order reductions -> bit-identical
    best_w2, xy_w2, tab_w2 = hb.shift_cPSNR_argmax(*args)
    assert np.array_equal(tab_w2.cpu().numpy(), tab_w, equal_nan=True) and torch.equal(xy_w2, xy_w)


@pytest.mark.parametrize("kind", ["bias_dominates", "local_bias", "exact_match", "nan_pixel", "inf_pixel", "huge_values",
                                  "masked_rows", "near_ties"])
def test_shift_cpsnr_onepass_hazards(hb, dev, kind):
    """The one-pass kernel computes sum(m d^2) - n b^2 from centred fp32 partial sums and hands the sites it does not
    trust (and everything that is not a positive finite number) to a two-pass fallback.  Inputs built to break exactly
    that: a brightness bias 3000 x the noise, a bias that changes across the image (each work item centres on its own
    first row), exact matches (cMSE = 0 -> +inf), NaN / inf pixels, hr values far outside [0, 1], whole rows masked out
    (a work item may see an empty first row) and 49 nearly identical scores.  Against the oracle and the two-pass path."""
    rng = np.random.RandomState(len(kind) * 7 + 1)
    b, s = 4, 200                                                    # two column blocks (194 crop columns), several row bands
    sr = rng.rand(b, s, s).astype(np.float32)
    hm = (rng.rand(b, s, s) > 0.15).astype(np.float32)
    hr = np.roll(sr, (1, -2), (1, 2)).copy()
    if kind == "bias_dominates":
        hr = (hr + 0.3 + 1e-4 * rng.randn(b, s, s)).astype(np.float32)
    elif kind == "local_bias":
        ramp = np.linspace(-0.4, 0.4, s, dtype=np.float32)
        hr = (hr + ramp[None, :, None] + 0.2 * ramp[None, None, :] + 1e-3 * rng.randn(b, s, s)).astype(np.float32)
    elif kind == "exact_match":
        hr = (hr + 0.125).astype(np.float32)                         # exact in fp32: d - b == 0 at the matching site
    elif kind == "nan_pixel":
        hr = (hr + 0.01 * rng.randn(b, s, s)).astype(np.float32)
        hr[0, 50, 60] = np.nan                                       # (the reference asserts 0 <= sr <= 1, so only hr can carry one)
        hr[1, 100:103, 100] = np.nan
    elif kind == "inf_pixel":
        hr = (hr + 0.01 * rng.randn(b, s, s)).astype(np.float32)
        hr[0, 50, 60] = np.inf
        hm[1] = 0.0                                                  # and an empty map: 0 / 0
    elif kind == "huge_values":
        hr = (hr * 3000.0 + 500.0 + 0.5 * rng.randn(b, s, s)).astype(np.float32)
    elif kind == "masked_rows":
        hr = (hr + 0.05 + 0.01 * rng.randn(b, s, s)).astype(np.float32)
        hm[:, ::24] = 0.0
        hm[:, 1::24] = 0.0
        hm[2, :120] = 0.0
    elif kind == "near_ties":
        hr = rng.rand(b, s, s).astype(np.float32)                    # uncorrelated: all 49 scores within ~1e-2 dB
    args = [torch.from_numpy(a).to(dev) for a in (sr, hr, hm)]
    with np.errstate(all="ignore"):
        ref_max, ref_arg, ref_sites = scoring_oracle.shift_cpsnr(sr, hr, hm)
    ref_sites = ref_sites.T
    best, xy, tab = hb.shift_cPSNR_argmax(*args)
    hb.scoring_debug_set("cpsnr_onepass", 0)
    try:
        best2, xy2, tab2 = hb.shift_cPSNR_argmax(*args)
    finally:
        hb.scoring_debug_set("cpsnr_onepass", 1)
    for t, a in ((tab, xy), (tab2, xy2)):
        t = t.cpu().numpy()
        assert np.array_equal(np.isnan(t), np.isnan(ref_sites))
        assert np.array_equal(np.isposinf(t), np.isposinf(ref_sites))
        fin = np.isfinite(ref_sites)
        assert np.abs(t[fin] - ref_sites[fin]).max(initial=0.0) <= CPSNR_KERNEL_GATE_DB
        assert np.array_equal((a[:, 0] * 7 + a[:, 1]).cpu().numpy(), ref_arg)
    assert torch.equal(torch.isnan(best), torch.isnan(best2))


def test_shift_cpsnr_onepass_differential(hb, dev):
    """Random sizes (crop widths that 6 does and does not divide, one to four column blocks, one to many row bands), batch
    sizes, maps (dense, sparse, striped, empty) and brightness offsets / noise levels over four decades: the one-pass search
    against the two-pass window kernels -- tables within 1e-4 dB, same NaN / inf pattern, and the best shift may differ only
    where the two-pass table itself has the two candidates within 2e-4 dB (tools/cpsnr_fuzz.py is the long version)."""
    rng = np.random.RandomState(4321)
    for case in range(36):
        s = int(rng.choice([8, 12, 20, 36, 64, 100, 128, 132, 200, 260, 384, 388]))
        b = int(rng.randint(1, 5)) if s > 200 else int(rng.randint(1, 24))
        sr = rng.rand(b, s, s).astype(np.float32)
        kind = int(rng.randint(0, 6))
        hr = np.roll(sr, (int(rng.randint(-3, 4)), int(rng.randint(-3, 4))), (1, 2)).copy() if kind != 0 else rng.rand(b, s, s).astype(np.float32)
        hr = hr + np.float32(rng.uniform(-0.3, 0.3)) + np.float32(10.0 ** rng.uniform(-5, -1)) * rng.randn(b, s, s).astype(np.float32)
        if kind == 2:
            hr = hr + np.linspace(-0.2, 0.2, s, dtype=np.float32)[None, :, None]
        hm = (rng.rand(b, s, s) > rng.uniform(0.0, 0.9)).astype(np.float32)
        if kind == 4:
            hm[:, ::int(rng.randint(2, 9))] = 0.0
        if kind == 5:
            hm[rng.randint(0, b)] = 0.0
        args = [torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(dev) for a in (sr, hr, hm)]
        _, xy1, tab1 = hb.shift_cPSNR_argmax(*args)
        hb.scoring_debug_set("cpsnr_onepass", 0)
        try:
            _, xy2, tab2 = hb.shift_cPSNR_argmax(*args)
        finally:
            hb.scoring_debug_set("cpsnr_onepass", 1)
        t1, t2 = tab1.cpu().numpy(), tab2.cpu().numpy()
        assert np.array_equal(np.isnan(t1), np.isnan(t2)) and np.array_equal(np.isinf(t1), np.isinf(t2)), (case, b, s, kind)
        fin = np.isfinite(t2)
        assert np.abs(t1[fin] - t2[fin]).max(initial=0.0) <= CPSNR_KERNEL_GATE_DB, (case, b, s, kind)
        a1 = (xy1[:, 0] * 7 + xy1[:, 1]).cpu().numpy()
        a2 = (xy2[:, 0] * 7 + xy2[:, 1]).cpu().numpy()
        for i in range(b):
            assert a1[i] == a2[i] or abs(t2[i, a1[i]] - t2[i, a2[i]]) <= 2e-4, (case, b, s, kind, i)


def test_shift_cpsnr_rejects_bad_arguments(hb, dev):
    sr = torch.rand(1, 20, 24, device=dev)
    with pytest.raises(RuntimeError):
        hb.shift_cPSNR(sr, sr, sr)                                              # non-square
    sq = torch.rand(1, 20, 20, device=dev)
    with pytest.raises(RuntimeError):
        hb.shift_cPSNR(sq, sq, sq, border_w=10)                                 # nothing left of a 20 x 20 image


@pytest.mark.parametrize("border_w,s", [(4, 24), (5, 40), (8, 33)])
def test_shift_cpsnr_any_border_matches_oracle(hb, dev, border_w, s):
    """Evaluator.py:52 accepts any border_w; above 3 the search takes the one-block-per-site kernels."""
    rng = np.random.RandomState(100 + border_w)
    b = 3
    sr = rng.rand(b, s, s).astype(np.float32)
    hr = np.clip(np.roll(sr, (2, -1), (1, 2)) + 0.01 * rng.randn(b, s, s), 0, 1).astype(np.float32)
    hm = (rng.rand(b, s, s) > 0.2).astype(np.float32)
    best, xy, tab = hb.shift_cPSNR_argmax(torch.from_numpy(sr).to(dev), torch.from_numpy(hr).to(dev), torch.from_numpy(hm).to(dev),
                                          border_w=border_w)
    span = 2 * border_w + 1
    for i in range(b):
        ref_max, ref_arg, ref_sites = scoring_oracle.shift_cpsnr(sr[i], hr[i], hm[i], border_w=border_w)
        assert abs(float(best[i]) - float(ref_max)) <= CPSNR_KERNEL_GATE_DB
        assert int(xy[i, 0]) * span + int(xy[i, 1]) == int(ref_arg)
        assert np.abs(tab[i].cpu().numpy() - np.asarray(ref_sites).T.reshape(-1)).max() <= CPSNR_KERNEL_GATE_DB


# ---------------------------------------------------------------------------- the composite path (C4)
def test_full_scoring_path_against_oracle(hb, net, dev):
    """HRNet -> lanczos_shift -> clip -> shift_cPSNR on 16-view imagesets (BASELINE configs[3], reduced batch).
    HR is built from the ORACLE's fp32 SR (roll + bias + sigma = 0.01 noise) so the best
    shift is known and well separated; end-to-end gate: cPSNR within 0.01 dB, identical argmax."""
    b, l, s = 3, 16, 64
    rng = np.random.RandomState(17)
    lrs = rng.rand(b, l, s, s).astype(np.float32)
    alphas = np.ones((b, l), dtype=np.float32)
    params = hrnet_oracle.make_params(cases.WEIGHT_SEED)
    sr_ref = hrnet_oracle.hrnet_forward(params, lrs, alphas).numpy()[:, 0]
    shift = rng.uniform(-1, 1, size=(b, 2)).astype(np.float32)
    moved_ref = np.clip(scoring_oracle.lanczos_shift(sr_ref[None], shift, p=5)[0], 0, 1)
    rolls = rng.randint(-3, 4, size=(b, 2))
    hr = np.stack([np.roll(moved_ref[i], tuple(rolls[i]), (0, 1)) for i in range(b)])
    hr = np.clip(hr + 0.02 + 0.01 * rng.randn(*hr.shape), 0, 1).astype(np.float32)
    hm = (rng.rand(*hr.shape) > 0.1).astype(np.float32)
    ref_scores = [scoring_oracle.shift_cpsnr(moved_ref[i], hr[i], hm[i]) for i in range(b)]

    sr = net(torch.from_numpy(lrs).to(dev), torch.from_numpy(alphas).to(dev))[:, 0]          # (B, 3s, 3s)
    moved = hb.lanczos_shift(sr[None], torch.from_numpy(shift).to(dev), p=5, a=3, N=7)[0]    # ShiftNet.py:87-89 layout
    best, xy, _ = hb.shift_cPSNR_argmax(moved, torch.from_numpy(hr).to(dev), torch.from_numpy(hm).to(dev),
                                        clip_sr=True)
    best, xy = best.cpu().numpy(), xy.cpu().numpy()
    for i in range(b):
        assert abs(best[i] - ref_scores[i][0]) <= CPSNR_GATE_DB
        assert xy[i, 0] * 7 + xy[i, 1] == ref_scores[i][1] == (3 + rolls[i, 0]) * 7 + (3 + rolls[i, 1])


def test_c4_full_batch_scoring_path(hb, net, dev):
    """BASELINE.json configs[3] at FULL batch: 32 imagesets of 16 views, 128^2 -> 384^2, HRNet -> lanczos_shift -> clip ->
    shift_cPSNR(border 3).  The oracle runs the whole chain for a sample of the batch (CPU time); for every imageset
    the known roll must be the winning shift (HR is built from the device SR for the others).
    Gates: SR 1e-2 max-abs, cPSNR 0.01 dB, argmax exact (sigma = 0.01 construction)."""
    b, l, s = 32, 16, 128
    rng = np.random.RandomState(41)
    lrs = rng.rand(b, l, s, s).astype(np.float32)
    alphas = np.ones((b, l), dtype=np.float32)
    shift = rng.uniform(-1, 1, size=(b, 2)).astype(np.float32)
    rolls = rng.randint(-3, 4, size=(b, 2))
    params = hrnet_oracle.make_params(cases.WEIGHT_SEED)
    sample = [0, 17, 31]
    sr_ref = hrnet_oracle.hrnet_forward(params, lrs[sample], alphas[sample]).numpy()[:, 0]
    moved_ref = np.clip(scoring_oracle.lanczos_shift(sr_ref[None], shift[sample], p=5)[0], 0, 1)

    sr = net(torch.from_numpy(lrs).to(dev), torch.from_numpy(alphas).to(dev))[:, 0]
    assert np.abs(sr[sample].cpu().numpy() - sr_ref).max() <= SR_REGRESSION_GATE
    moved = hb.lanczos_shift(sr[None], torch.from_numpy(shift).to(dev), p=5, a=3, N=7)[0]
    assert np.abs(moved[sample].cpu().numpy().clip(0, 1) - moved_ref).max() <= SR_REGRESSION_GATE
    # HR: from the oracle's chain for the sampled imagesets, from the device chain for the rest
    base = moved.clamp(0, 1).cpu().numpy()
    for k, i in enumerate(sample):
        base[i] = moved_ref[k]
    hr = np.stack([np.roll(base[i], tuple(rolls[i]), (0, 1)) for i in range(b)])
    hr = np.clip(hr + 0.02 + 0.01 * rng.randn(*hr.shape), 0, 1).astype(np.float32)
    hm = (rng.rand(*hr.shape) > 0.1).astype(np.float32)
    best, xy, _ = hb.shift_cPSNR_argmax(moved, torch.from_numpy(hr).to(dev), torch.from_numpy(hm).to(dev), clip_sr=True)
    best, xy = best.cpu().numpy(), xy.cpu().numpy()
    assert np.array_equal(xy[:, 0] * 7 + xy[:, 1], (3 + rolls[:, 0]) * 7 + (3 + rolls[:, 1]))
    for k, i in enumerate(sample):
        ref_max, ref_arg, _ = scoring_oracle.shift_cpsnr(moved_ref[k], hr[i], hm[i])
        assert abs(best[i] - ref_max) <= CPSNR_GATE_DB, (best[i], ref_max)
        assert xy[i, 0] * 7 + xy[i, 1] == ref_arg
    # kernel parity on the device tensors themselves (same SR in, strict gate) for two more imagesets
    moved_np = moved.clamp(0, 1).cpu().numpy()
    for i in (5, 23):
        ref_max, ref_arg, _ = scoring_oracle.shift_cpsnr(moved_np[i], hr[i], hm[i])
        assert abs(best[i] - ref_max) <= CPSNR_KERNEL_GATE_DB and xy[i, 0] * 7 + xy[i, 1] == ref_arg


def test_c3_shard_shape_one_imageset_vs_oracle(net, dev):
    """BASELINE.json configs[2], the per-rank shard: 32 imagesets x 32 views of 128 x 128 (five fusion levels, 10 GiB of
    workspace).  One imageset of the full batch against the oracle, plus batch independence at that size."""
    b, l, s = 32, 32, 128
    g = torch.Generator().manual_seed(32)
    lrs = torch.rand(b, l, s, s, generator=g)
    alphas = torch.ones(b, l)
    tl, ta = lrs.to(dev), alphas.to(dev)
    sr = net(tl, ta)
    assert tuple(sr.shape) == (b, 1, 3 * s, 3 * s) and torch.isfinite(sr).all()
    params = hrnet_oracle.make_params(cases.WEIGHT_SEED)
    ref = hrnet_oracle.hrnet_forward(params, lrs[7:8].numpy(), alphas[7:8].numpy()).numpy()
    err = np.abs(sr[7:8].cpu().numpy() - ref).max()
    assert err <= SR_GATE and err <= 2e-3, err                 # five levels of bf16 rounding: 1.1e-3 measured
    assert torch.equal(net(tl[7:8], ta[7:8])[0], sr[7])


def test_c5_large_tile_one_imageset_vs_oracle(net, dev):
    """BASELINE.json configs[4]: 512 x 512 LR (four 128-pixel column tiles per row), 8 views, batch 8 -> 1536 x 1536."""
    b, l, s = 8, 8, 512
    g = torch.Generator().manual_seed(512)
    lrs = torch.rand(b, l, s, s, generator=g)
    alphas = torch.ones(b, l)
    tl, ta = lrs.to(dev), alphas.to(dev)
    sr = net(tl, ta)
    assert tuple(sr.shape) == (b, 1, 3 * s, 3 * s) and torch.isfinite(sr).all()
    params = hrnet_oracle.make_params(cases.WEIGHT_SEED)
    ref = hrnet_oracle.hrnet_forward(params, lrs[3:4].numpy(), alphas[3:4].numpy()).numpy()
    err = np.abs(sr[3:4].cpu().numpy() - ref).max()
    assert err <= SR_GATE and err <= SR_REGRESSION_GATE, err
    assert torch.equal(net(tl[3:4], ta[3:4])[0], sr[3])


# ---------------------------------------------------------------------------- the caller
@pytest.mark.parametrize("name", list(cases.PREDICT_CASES))
def test_get_sr_and_score_matches_reference_golden(net, golden, name):
    """predict.get_sr_and_score (predict.py:17-49): collate to min_L = 16, forward, clip, shifted cPSNR."""
    from highres_net_b200.predict import get_sr_and_score, get_sr_and_score_batch
    g = golden["predict"]
    n, s, has_hr = cases.PREDICT_CASES[name]
    imset = {"name": name, "lr": torch.from_numpy(cases.predict_lrs(name)),
             "hr": torch.from_numpy(g[name + "__hr"]) if has_hr else None,
             "hr_map": torch.from_numpy(g[name + "__hr_map"]) if has_hr else torch.ones(3 * s, 3 * s)}
    sr, score = get_sr_and_score(imset, net, min_L=cases.PREDICT_MIN_L)
    assert isinstance(sr, np.ndarray) and sr.shape == (3 * s, 3 * s)
    assert np.abs(sr - g[name + "__sr"]).max() <= SR_REGRESSION_GATE
    if has_hr:
        assert abs(float(score) - float(g[name + "__score"])) <= CPSNR_GATE_DB
        srs, scores = get_sr_and_score_batch([imset, imset], net, min_L=cases.PREDICT_MIN_L)
        assert srs.shape == (2, 3 * s, 3 * s) and np.array_equal(srs[0], sr) and np.array_equal(srs[1], sr)
        assert scores.shape == (2,) and abs(float(scores[1]) - float(score)) <= 1e-5
    else:
        assert score is None


def _evaluate_datasets(golden):
    g = golden["evaluate"]
    sets = {"train": [], "val": [], "test": []}
    for name, (case, split) in cases.EVALUATE_SETS.items():
        n, s, has_hr = cases.PREDICT_CASES[case]
        sets[split].append({"name": name, "lr": torch.from_numpy(cases.evaluate_lrs(name)),
                            "hr": torch.from_numpy(g[name + "__hr"]) if has_hr else None,
                            "hr_map": torch.from_numpy(g[name + "__hr_map"]) if has_hr else torch.ones(3 * s, 3 * s),
                            "clearances": cases.evaluate_clearances(name)})
    return sets


@pytest.mark.parametrize("batch_size", [1, 2, 32])
def test_evaluate_matches_reference_golden(net, golden, batch_size):
    """predict.evaluate (predict.py:103-135): same three dicts as the reference, scores within the gate; batching
    imagesets of equal shape does not change a single score (bit-identical to one imageset per forward)."""
    from highres_net_b200.predict import evaluate, get_sr_and_score
    g = golden["evaluate"]
    sets = _evaluate_datasets(golden)
    scores, clerances, part = evaluate(net, sets["train"], sets["val"], sets["test"], min_L=cases.PREDICT_MIN_L,
                                       batch_size=batch_size)
    assert set(scores) == set(cases.EVALUATE_SETS) == set(clerances) == set(part)
    for name, (case, split) in cases.EVALUATE_SETS.items():
        assert part[name] == split
        assert np.array_equal(clerances[name], cases.evaluate_clearances(name))
        if cases.PREDICT_CASES[case][2]:
            assert abs(float(scores[name]) - float(g[name + "__score"])) <= CPSNR_GATE_DB
            one = next(im for im in sets[split] if im["name"] == name)
            assert np.float32(scores[name]) == get_sr_and_score(one, net, min_L=cases.PREDICT_MIN_L)[1]
        else:
            assert scores[name] is None


def test_model_wrapper_and_load_model(hb, golden, dev, tmp_path):
    """predict.Model / load_model (predict.py:83-100, 200-217): a checkpoint written the reference's way
    (torch.save(state_dict), train.py:220-222) loads into the B200 module and scores like the reference."""
    from highres_net_b200.predict import Model
    g = golden["evaluate"]
    ckpt = tmp_path / "HRNet.pth"
    torch.save(hrnet_oracle.make_params(cases.WEIGHT_SEED), ckpt)
    config = {"network": hrnet_oracle.DEFAULT_NETWORK_CONFIG, "training": {"min_L": cases.PREDICT_MIN_L}}
    m = Model(config)
    m.load_checkpoint(str(ckpt))
    sets = _evaluate_datasets(golden)
    sr, score = m(sets["train"][0])
    assert sr.shape == (96, 96) and abs(float(score) - float(g["imgset_a__score"])) <= CPSNR_GATE_DB
    scored = [n for n, (c, _) in cases.EVALUATE_SETS.items() if cases.PREDICT_CASES[c][2]]
    table = m.evaluate(sets["train"], sets["val"], [], {n: cases.EVALUATE_BASELINE[n] for n in scored})
    assert list(table.index) == [str(n) for n in g["benchmark__index"]]
    assert np.abs(table["model"].to_numpy(dtype=np.float64) - g["benchmark__model"]).max() <= CPSNR_GATE_DB
    assert np.allclose(table["score"].to_numpy(dtype=np.float64), g["benchmark__score"], rtol=1e-3)


def test_img_as_uint_matches_the_skimage_rule(dev):
    """predict.py:176 (img_as_uint before the PNG writer) on the device: every uint16 value survives the round trip
    through img_as_float_u16, half-way cases round to even, negatives clip to 0, values outside [-1, 1] raise."""
    from highres_net_b200 import predict
    from oracle import predict_oracle
    codes = torch.arange(65536, dtype=torch.int32).to(torch.uint16).to(dev)
    back = predict.img_as_uint_u16(predict.img_as_float_u16(codes))
    assert torch.equal(back.cpu().view(torch.int16), codes.cpu().view(torch.int16))
    rng = np.random.RandomState(9)
    x = np.concatenate([rng.uniform(-1, 1, 100000), (np.arange(0, 2000) + 0.5) / 65535.0, [-1.0, 1.0, -0.0, 0.0, 1e-9]]).astype(np.float32)
    got = predict.img_as_uint_u16(torch.from_numpy(x).to(dev)).cpu().view(torch.int16).numpy().view(np.uint16)
    assert np.array_equal(got, predict_oracle.img_as_uint(x))
    with pytest.raises(ValueError):
        predict.img_as_uint_u16(torch.tensor([0.5, 1.001], device=dev))
    with pytest.raises(RuntimeError):
        predict.img_as_uint_u16(torch.tensor([0.5]))                               # host tensor: no CPU fallback


# ---------------------------------------------------------------------------- train.get_loss twin
LOSS_REL_GATE = 1e-5      # fp32 element ops as in the reference, fp64 sums


@pytest.mark.parametrize("name", list(cases.LOSS_CASES))
@pytest.mark.parametrize("metric", cases.LOSS_METRICS)
def test_get_loss_matches_reference_golden(hb, dev, golden, name, metric):
    sr, hr, hm = (torch.from_numpy(x).to(dev) for x in cases.loss_inputs(name))
    got = hb.get_loss(sr, hr, hm, metric=metric)
    assert got.shape == (sr.shape[0],) and got.dtype == torch.float32 and got.is_cuda
    ref = golden["loss"][f"{name}__{metric}"]
    assert np.abs(got.cpu().numpy() / ref - 1).max() <= LOSS_REL_GATE
    assert np.abs(got.cpu().numpy() / scoring_oracle.clear_loss(*cases.loss_inputs(name), metric) - 1).max() <= LOSS_REL_GATE


def test_get_loss_agrees_with_evaluator_on_binary_masks(hb, dev):
    """For a 0/1 mask m == m^2, so train.get_loss('cPSNR') equals Evaluator.cPSNR (Evaluator.py:15-37) image by image."""
    sr, hr, hm = (torch.from_numpy(x).to(dev) for x in cases.loss_inputs("full_2_384"))
    loss = hb.get_loss(sr, hr, hm, metric="cPSNR").cpu().numpy()
    for i in range(sr.shape[0]):
        assert abs(loss[i] - float(hb.cPSNR(sr[i], hr[i], hm[i]))) <= 1e-4
    again = hb.get_loss(sr, hr, hm, metric="cPSNR").cpu().numpy()
    assert np.array_equal(loss, again)                                          # fixed-order reduction: deterministic


def test_get_loss_degenerate_and_errors(hb, dev):
    sr = torch.rand(2, 24, 24, device=dev)
    zero = torch.zeros_like(sr)
    assert torch.isnan(hb.get_loss(sr, sr * 0.5, zero, metric="cMSE")).all()    # 0 / 0 like the reference
    assert torch.isinf(hb.get_loss(sr, sr, torch.ones_like(sr), metric="cPSNR")).all()
    assert (hb.get_loss(sr, sr, torch.ones_like(sr), metric="masked_MSE") == 0).all()
    with pytest.raises(Exception):
        hb.get_loss(sr.cpu(), sr.cpu(), zero.cpu())                             # no CPU fallback


def test_get_crop_mask_matches_reference_golden(hb, golden):
    for ps, cs in ((32, 3), (4, 1), (64, 6)):
        assert np.array_equal(hb.get_crop_mask(ps, cs).numpy(), golden["loss"][f"crop_{ps}_{cs}"])


# ---------------------------------------------------------------------------- train.apply_shifts / ShiftNet.transform (8a10)
@pytest.mark.parametrize("name", list(cases.APPLY_SHIFTS_CASES))
def test_apply_shifts_matches_reference_golden(hb, dev, golden, name):
    images, thetas = cases.apply_shifts_inputs(name)
    out = hb.apply_shifts(None, torch.from_numpy(images).to(dev), torch.from_numpy(thetas).to(dev), dev)
    assert out.shape == images.shape and out.is_cuda
    assert np.abs(out.cpu().numpy() - golden["apply_shifts"][name]).max() <= LANCZOS_GATE


def test_apply_shifts_integer_theta_is_a_roll(hb, dev):
    """theta = (dx, dy) = (1, -2): every interior pixel moves by exactly that many pixels (taps collapse to a delta)."""
    img = torch.rand(1, 2, 32, 32, device=dev)
    theta = torch.tensor([[[1.0, -2.0], [0.0, 0.0]]], device=dev)
    out = hb.apply_shifts(None, img, theta, dev)
    assert torch.allclose(out[0, 1], img[0, 1], atol=2e-6)
    ref = scoring_oracle.apply_shifts(img.cpu().numpy(), theta.cpu().numpy())
    assert np.abs(out.cpu().numpy() - ref).max() <= LANCZOS_GATE


# ---------------------------------------------------------------------------- 16-bit view ingestion
def test_u16_views_scale_like_the_dataloader(hb, dev):
    """DataLoader.py:195-198 = skimage.img_as_float(uint16).astype(float32): x / 65535 (or x * (1 / 65535)) in float64,
    rounded once to float32.  All 65536 inputs must come out bit-identical."""
    from highres_net_b200 import predict
    x = np.arange(65536, dtype=np.uint16)
    got = predict.img_as_float_u16(torch.from_numpy(x).to(dev)).cpu().numpy()
    assert np.array_equal(got, (x.astype(np.float64) / 65535.0).astype(np.float32))
    assert np.array_equal(got, (x.astype(np.float64) * (1.0 / 65535.0)).astype(np.float32))
    with pytest.raises(Exception):
        predict.img_as_float_u16(torch.from_numpy(x))                          # host tensor: no CPU fallback


def test_forward_host_accepts_raw_u16_views(net, dev):
    rng = np.random.RandomState(8)
    raw = rng.randint(0, 1 << 14, size=(3, 5, 32, 32)).astype(np.uint16)       # Proba-V views are 14-bit in 16-bit PNGs
    alphas = np.ones((3, 5), dtype=np.float32)
    alphas[1, 3:] = 0
    raw[1, 3:] = 0
    as_float = torch.from_numpy((raw.astype(np.float64) / 65535.0).astype(np.float32))
    a = net.forward_host(torch.from_numpy(raw), torch.from_numpy(alphas), device=dev)
    b = net.forward_host(as_float, torch.from_numpy(alphas), device=dev)
    assert torch.equal(a, b)


# ---------------------------------------------------------------------------- ShiftNet
# bf16 activations / weights with fp32 accumulation through 8 conv layers and a K = 32768 GEMM, against the fp32 reference
SHIFTNET_GATE = 1e-2      # absolute, on thetas of magnitude ~0.2-0.7: bf16 activations and weights through 8 conv layers and a K = 32768
                          # contraction; measured 4.6e-3, so the gate keeps 2x of headroom instead of 4x


@pytest.fixture(scope="module")
def shiftnet(hb, dev):
    from oracle import shiftnet_oracle
    net = hb.ShiftNet().eval()
    net.load_state_dict(shiftnet_oracle.make_params(0), strict=True)
    return net.to(dev)


def test_shiftnet_forward_matches_reference_golden(shiftnet, golden, dev):
    from oracle import shiftnet_oracle
    g = golden["shiftnet"]
    x = torch.from_numpy(shiftnet_oracle.make_pairs(6, 0)).to(dev)
    theta = shiftnet(x)
    assert theta.shape == (6, 2) and theta.dtype == torch.float32 and theta.is_cuda
    err = np.abs(theta.cpu().numpy() - g["theta"]).max()
    assert err <= SHIFTNET_GATE, err
    # the differences BETWEEN pairs (what registration is about) survive the bf16 path
    d_ref = g["theta"] - g["theta"].mean(0)
    d_got = theta.cpu().numpy() - theta.cpu().numpy().mean(0)
    assert np.abs(d_got - d_ref).max() <= SHIFTNET_GATE
    assert torch.equal(shiftnet(x), theta)                                        # deterministic (fixed-order split-K)
    # batch independence: a pair gives the same theta whatever else is in the batch (bit-exact), also across row tiles
    big = torch.from_numpy(shiftnet_oracle.make_pairs(150, 2)).to(dev)
    t_big = shiftnet(big)
    # the 32 x 32 and 16 x 16 layers put 3 / 7 images into one MMA tile: same thetas as one image row per tile
    for n in (150, 5, 1):
        shiftnet.debug_set(dev, "img_group", 0)
        plain = shiftnet(big[:n])
        shiftnet.debug_set(dev, "img_group", 1)
        assert torch.equal(plain, t_big[:n]), n
    # MaxPool2d(2) inside the conv epilogue (row pairs + a lane shuffle) against the stand-alone pooling launch
    for n in (150, 5, 1):
        for grouped in (1, 0):
            shiftnet.debug_set(dev, "img_group", grouped)
            shiftnet.debug_set(dev, "fused_pool", 0)
            unfused = shiftnet(big[:n])
            shiftnet.debug_set(dev, "fused_pool", 1)
            assert torch.equal(unfused, shiftnet(big[:n])), (n, grouped)
            assert torch.equal(unfused, t_big[:n]), (n, grouped)
    shiftnet.debug_set(dev, "img_group", 1)
    assert torch.equal(shiftnet(big[140:145]), t_big[140:145])
    assert torch.equal(shiftnet(big[:1]), t_big[:1])


def test_shiftnet_register_batch_matches_reference_golden(hb, shiftnet, golden, dev):
    from oracle import shiftnet_oracle
    g = golden["shiftnet"]
    pairs = shiftnet_oracle.make_pairs(6, 1).reshape(2, 3, 2, 128, 128)
    lrs = torch.from_numpy(np.ascontiguousarray(pairs[:, :, 1])).to(dev)
    reference = torch.from_numpy(np.ascontiguousarray(pairs[:, 0, 0][:, None])).to(dev)
    thetas = hb.register_batch(shiftnet, lrs, reference)
    assert thetas.shape == (2, 3, 2)
    assert np.abs(thetas.cpu().numpy() - g["register_thetas"]).max() <= SHIFTNET_GATE
    # thetas feed transform / apply_shifts (train.py:47-63) like in the reference
    moved = hb.apply_shifts(shiftnet, lrs, thetas, dev)
    assert moved.shape == lrs.shape


def test_shiftnet_weight_updates_and_error_paths(hb, dev):
    from oracle import shiftnet_oracle
    net = hb.ShiftNet().to(dev)
    x = torch.from_numpy(shiftnet_oracle.make_pairs(2, 3)).to(dev)
    with pytest.raises(RuntimeError):
        net(x)                                                                    # training mode: not supported
    net.eval()
    assert float(net(x).abs().max()) == 0.0                                       # fc2 starts at zero (ShiftNet.py:48)
    net.load_state_dict(shiftnet_oracle.make_params(0), strict=True)              # new weights must reach the device copy
    ref = shiftnet_oracle.shiftnet_forward(shiftnet_oracle.make_params(0), x.cpu().numpy()).numpy()
    assert np.abs(net(x).cpu().numpy() - ref).max() <= SHIFTNET_GATE
    with pytest.raises(RuntimeError):
        net(torch.zeros(1, 2, 64, 64, device=dev))                                # fc1 needs 128 x 128 crops
    with pytest.raises(ValueError):
        net(torch.zeros(1, 3, 128, 128, device=dev))


# ---------------------------------------------------------------------------- several GPUs driven by ONE process
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two visible GPUs")
def test_one_process_two_devices(hb):
    """One handle per device: the same module / functions used on cuda:0 and cuda:1 from one process give identical
    results (per-device kernel attributes, memory pools, handles and streams)."""
    from oracle import shiftnet_oracle
    d0, d1 = torch.device("cuda:0"), torch.device("cuda:1")
    g = torch.Generator().manual_seed(77)
    lrs, al = torch.rand(3, 4, 64, 64, generator=g), torch.ones(3, 4)
    pairs = torch.from_numpy(shiftnet_oracle.make_pairs(5, 4))
    sr_, hr_, hm_ = torch.rand(2, 96, 96, generator=g), torch.rand(2, 96, 96, generator=g), (torch.rand(2, 96, 96, generator=g) > 0.1).float()
    outs = []
    for dev in (d1, d0, d1):                       # cuda:1 first: nothing may rely on device 0 having been set up
        net = hb.HRNet(hrnet_oracle.DEFAULT_NETWORK_CONFIG).eval()
        net.load_state_dict(hrnet_oracle.make_params(cases.WEIGHT_SEED))
        net = net.to(dev)
        sn = hb.ShiftNet().eval()
        sn.load_state_dict(shiftnet_oracle.make_params(0))
        sn = sn.to(dev)
        sr = net(lrs.to(dev), al.to(dev))
        host = net.forward_host(lrs.pin_memory(), al.pin_memory(), device=dev)
        theta = sn(pairs.to(dev))
        moved = hb.lanczos_shift(sr_.to(dev)[None], torch.tensor([[0.3, -0.6], [1.0, 0.0]], device=dev), p=5)
        best, xy, tab = hb.shift_cPSNR_argmax(sr_.to(dev), hr_.to(dev), hm_.to(dev))
        assert sr.device == dev and theta.device == dev and tab.device == dev
        outs.append([t.cpu() for t in (sr, host, theta, moved, tab, xy)])
    for o in outs[1:]:
        for a, b in zip(o, outs[0]):
            assert torch.equal(a, b)
    # the C entry points run on the handle's device and restore the caller's current device
    import ctypes
    from highres_net_b200 import _lib
    torch.cuda.set_device(0)
    net = hb.HRNet(hrnet_oracle.DEFAULT_NETWORK_CONFIG).eval().to(d1)
    handle = net._handle_for(d1)                                   # hrn_create + hrn_set_weight on device 1
    assert torch.cuda.current_device() == 0
    x1, a1 = lrs.to(d1), al.to(d1)
    out1 = torch.empty(3, 1, 192, 192, device=d1)
    torch.cuda.synchronize(d1)
    _lib.check(_lib.load().hrn_forward(handle, x1.data_ptr(), a1.data_ptr(), 3, 4, 64, 64, out1.data_ptr(), None), "hrn_forward")
    assert torch.cuda.current_device() == 0
    torch.cuda.synchronize(d1)
    assert torch.equal(out1.cpu(), net(x1, a1).cpu())


# ---------------------------------------------------------------------------- the training-step forward, end to end
def test_trainstep_forward_value_through_every_drop_in(hb, shiftnet, golden, dev):
    """train.py:174-187 without autograd: fusion model -> register_batch on the 128 x 128 centre crops -> apply_shifts ->
    -get_loss('cPSNR') with get_crop_mask -> + lambda * mean(shifts)^2, every piece being the B200 drop-in, against the
    value the unmodified reference computes (oracle/make_golden_trainstep.py)."""
    from oracle.make_golden_trainstep import inputs
    g = golden["trainstep"]
    lrs, alphas = inputs()
    fusion = hb.HRNet(hrnet_oracle.DEFAULT_NETWORK_CONFIG).eval()
    fusion.load_state_dict(hrnet_oracle.make_params(cases.WEIGHT_SEED))
    fusion = fusion.to(dev)
    hrs, hr_maps = torch.from_numpy(g["hr"]).to(dev), torch.from_numpy(g["hr_map"]).to(dev)
    off, patch = int(g["offset"]), 64
    torch_mask = hb.get_crop_mask(patch_size=patch, crop_size=int(g["crop"]))
    srs = fusion(torch.from_numpy(lrs).to(dev), torch.from_numpy(alphas).to(dev))                       # train.py:174
    assert np.abs(srs.cpu().numpy() - g["srs"]).max() <= SR_REGRESSION_GATE
    shifts = hb.register_batch(shiftnet, srs[:, :, off:off + 128, off:off + 128].contiguous(),
                               reference=hrs[:, off:off + 128, off:off + 128].reshape(-1, 1, 128, 128))   # :177-179
    assert np.abs(shifts.cpu().numpy() - g["shifts"]).max() <= SHIFTNET_GATE
    srs_shifted = hb.apply_shifts(shiftnet, srs, shifts, dev)[:, 0]                                      # :180
    assert np.abs(srs_shifted.cpu().numpy() - g["srs_shifted"]).max() <= 5e-3
    cropped_mask = torch_mask[0].to(dev) * hr_maps                                                       # :183
    loss = -hb.get_loss(srs_shifted, hrs, cropped_mask, metric="cPSNR")                                  # :185
    assert np.abs(loss.cpu().numpy() - g["loss"]).max() <= 0.05                                          # dB
    total = loss.mean() + float(g["lam"]) * shifts.mean() ** 2                                           # :186-187
    assert abs(float(total) - float(g["total"])) <= 0.05


# ---------------------------------------------------------------------------- N > 1: the design's only collective on NCCL
def _nccl_worker(rank, world, port, results):
    import os
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    import highres_net_b200 as hb
    from highres_net_b200 import distributed as hd
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        model = hb.HRNet(hrnet_oracle.DEFAULT_NETWORK_CONFIG).eval()
        model.load_state_dict(hrnet_oracle.make_params(cases.WEIGHT_SEED))
        model = model.to(dev)
        n = 5                                                   # ragged: 3 + 2 imagesets
        g = torch.Generator().manual_seed(9)
        lrs = torch.rand(n, 4, 32, 32, generator=g).to(dev)
        alphas = torch.ones(n, 4, device=dev)
        sr_all = model(lrs, alphas)                             # every rank also computes the whole batch as the check
        hr = (torch.roll(sr_all[:, 0].clamp(0, 1), (1, -2), (1, 2)) + 0.02).clamp(0, 1)
        hm = (torch.rand(n, 96, 96, generator=g) > 0.1).float().to(dev)
        best_all, xy_all, _ = hb.shift_cPSNR_argmax(sr_all[:, 0], hr, hm, clip_sr=True)
        sr, best, xy = hd.sharded_forward_and_score(model, lrs, alphas, hr, hm)
        ok = torch.equal(sr, sr_all) and torch.equal(best, best_all) and torch.equal(xy, xy_all.to(torch.int64))
        sr1, best1, _ = hd.sharded_forward_and_score(model, lrs[:1], alphas[:1], hr[:1], hm[:1])   # fewer imagesets than ranks
        ok = ok and torch.equal(sr1, sr_all[:1]) and torch.equal(best1, best_all[:1])
        results[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two visible GPUs")
def test_sharded_forward_and_score_over_nccl():
    """Two ranks, one GPU each: batch shards through the real kernels, SR and (cPSNR, x, y) gathered with NCCL must be
    bit-identical to the single-GPU run (imagesets are independent and the kernels deterministic)."""
    import socket
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_nccl_worker, args=(2, port, results), nprocs=2, join=True)
    assert dict(results) == {0: True, 1: True}


# ---------------------------------------------------------------------------- N2 / N4: ragged collate on the device, file formats
@pytest.mark.parametrize("raw16", [False, True])
@pytest.mark.parametrize("min_l", [1, 4, 9, 16])
def test_collate_device_equals_host_collate(hb, dev, raw16, min_l):
    """hrn_collate (utils.py:63-113 on the device, only real views cross PCIe) against the host mirror: bit-exact
    padded batch and alphas, truncation at min_L, zero planes for the padding, uint16 views scaled like the DataLoader."""
    from highres_net_b200.predict import collateFunction, collate_device
    rng = np.random.RandomState(7 + min_l)
    batch = []
    for i, n in enumerate((3, 12, 9, 1, 16, 20)):
        lr = (rng.rand(n, 20, 20) * 65535).astype(np.uint16) if raw16 else rng.rand(n, 20, 20).astype(np.float32)
        batch.append({"name": f"s{i}", "lr": torch.from_numpy(lr), "hr": torch.rand(60, 60), "hr_map": torch.ones(60, 60)})
    ref_lrs, ref_alphas, ref_hr, ref_hm, ref_names = collateFunction(min_L=min_l)(batch)
    before = hb.kernel_launch_count()
    lrs, alphas, hrs, hms, names = collate_device(batch, min_l, dev)
    assert hb.kernel_launch_count() == before + 1 and lrs.is_cuda and alphas.is_cuda
    assert torch.equal(lrs.cpu(), ref_lrs) and torch.equal(alphas.cpu(), ref_alphas)
    assert torch.equal(hrs, ref_hr) and torch.equal(hms, ref_hm) and names == ref_names
    via_class = collateFunction(min_L=min_l, device=dev)(batch)
    assert torch.equal(via_class[0], lrs) and torch.equal(via_class[1], alphas)


def test_collate_device_odd_plane_size_and_errors(hb, dev):
    from highres_net_b200.predict import collateFunction, collate_device
    batch = [{"name": "a", "lr": torch.rand(2, 5, 7), "hr": None, "hr_map": None},
             {"name": "b", "lr": torch.rand(4, 5, 7), "hr": None, "hr_map": None}]
    lrs, alphas, hrs, hms, _ = collate_device(batch, 3, dev)                  # 35 pixels per plane: scalar path
    ref = collateFunction(min_L=3)(batch)
    assert torch.equal(lrs.cpu(), ref[0]) and torch.equal(alphas.cpu(), ref[1]) and hrs == [] and hms == [None, None]
    with pytest.raises(ValueError):
        collate_device(batch + [{"name": "c", "lr": torch.rand(2, 6, 6), "hr": None, "hr_map": None}], 3, dev)
    with pytest.raises(RuntimeError):
        collate_device(batch, 3, "cpu")


def test_disk_to_submission_through_every_io_end(hb, net, dev, tmp_path):
    """The whole N4 chain on the PIL-written fixture: save_clearance -> ImagesetDataset (native PNG decode, clearance
    order, raw uint16 views) -> collate on the device -> HRNet -> img_as_uint on the device -> native PNG encode ->
    stored submission.zip; each written image equals img_as_uint of the SR the float32 path gives."""
    import os
    import shutil
    import zipfile
    from highres_net_b200 import imageset_io as io
    from highres_net_b200.predict import get_sr_and_score, img_as_uint_u16
    src = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "imgsets")
    dirs = []
    for chan, name in (("RED", "imgset0001"), ("RED", "imgset0002"), ("NIR", "imgset0003")):
        dst = tmp_path / name
        shutil.copytree(os.path.join(src, chan, name), dst)
        dirs.append(str(dst))
    io.save_clearance_scores(dirs)
    cfg = {"create_patches": False, "patch_size": 8}
    raw_ds = io.ImagesetDataset(dirs, cfg, raw16=True)
    f32_ds = io.ImagesetDataset(dirs, cfg)
    out = str(tmp_path / "submission")
    archive = io.generate_submission_file(net, raw_ds, out=out, min_L=4, batch_size=2)
    with zipfile.ZipFile(archive) as z:
        assert z.testzip() is None and sorted(z.namelist()) == ["imgset0001.png", "imgset0002.png", "imgset0003.png"]
    for i, d in enumerate(dirs):
        name = os.path.basename(d)
        sr, score = get_sr_and_score(f32_ds[i], net, min_L=4)                 # the reference-shaped float32 route
        want = img_as_uint_u16(torch.from_numpy(sr).to(dev)).cpu().numpy()
        got = io.read_png_u16([os.path.join(out, name + ".png")], pin=False)[0].numpy()
        assert got.shape == (72, 72) and np.array_equal(got, want), name
        assert (score is None) == (name == "imgset0003")
