"""GPU bring-up battery (run on the B200 box):  python tools/bringup.py [--out gpurun_out/bringup.log]

Each group runs in its own subprocess under a timeout, so a trapped kernel in one
group cannot poison the CUDA context of the next.  Prints one line per check with
the measured error; it is a diagnostic tool, the pass/fail gates live in tests/.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def log(**kw):
    print(json.dumps(kw), flush=True)


def group_scoring():
    import numpy as np
    import torch
    import highres_net_b200 as hb
    from oracle import cases, scoring_oracle
    dev = torch.device("cuda:0")
    taps = hb.lanczos_kernel(torch.tensor(cases.LANCZOS_TAP_SHIFTS, dtype=torch.float32, device=dev).view(-1, 1))
    ref = scoring_oracle.lanczos_taps(np.array(cases.LANCZOS_TAP_SHIFTS, dtype=np.float32))
    log(check="lanczos_taps", max_err=float(np.abs(taps.cpu().numpy() - ref).max()))
    for name in cases.LANCZOS_CASES:
        img, shift, p = cases.lanczos_inputs(name)
        out = hb.lanczos_shift(torch.from_numpy(img).to(dev), torch.from_numpy(shift).to(dev), p=p).cpu().numpy()
        log(check="lanczos_shift", case=name, max_err=float(np.abs(out - scoring_oracle.lanczos_shift(img, shift, p=p)).max()))
    for name in cases.CPSNR_CASES:
        sr, hr, hm = cases.cpsnr_inputs(name)
        best, xy, table = hb.shift_cPSNR_argmax(sr, hr, hm)
        for i in range(sr.shape[0]):
            mx, am, sites = scoring_oracle.shift_cpsnr(sr[i], hr[i], hm[i])
            with np.errstate(invalid="ignore"):
                d = np.abs(table[i].astype(np.float64) - sites.astype(np.float64))
            d = d[np.isfinite(d)]
            log(check="shift_cpsnr", case=name, img=i, ref_max=float(mx), got_max=float(best[i]),
                ref_arg=int(am), got_arg=int(xy[i][0] * 7 + xy[i][1]), max_site_err_db=float(d.max()) if d.size else 0.0,
                nan_pattern_equal=bool(np.array_equal(np.isnan(table[i]), np.isnan(sites))))


def _layer_ref(x, w, b, slope):
    import torch
    import torch.nn.functional as F
    y = F.conv2d(x, w, b, padding=1)
    return F.prelu(y, slope) if slope is not None else y


def group_taps(max_ctas: int, size: int):
    """Isolate single taps of the first 64->64 tcgen05 conv (stage ENC(1))."""
    import numpy as np
    import torch
    import highres_net_b200 as hb
    from highres_net_b200 import hrnet as hm
    from oracle import hrnet_oracle
    dev = torch.device("cuda:0")
    params = hrnet_oracle.make_params(0)
    rng = np.random.RandomState(7)
    lrs = torch.from_numpy(rng.rand(1, 1, size, size).astype(np.float32)).to(dev)
    alphas = torch.ones(1, 1, device=dev)
    key = "encode.res_layers.0.block.0"
    full_w = params[key + ".weight"].clone()
    for taps in ([(1, 0)], [(1, 1)], [(1, 2)], [(0, 0)], [(2, 0)], [(0, 1), (2, 2)], None):
        w = full_w.clone()
        if taps is not None:
            mask = torch.zeros(3, 3)
            for ky, kx in taps:
                mask[ky, kx] = 1
            w = w * mask
        net = hb.HRNet(hrnet_oracle.DEFAULT_NETWORK_CONFIG).eval()
        sd = dict(params)
        sd[key + ".weight"] = w
        net.load_state_dict(sd)
        net = net.to(dev)
        net.debug_set(dev, "max_ctas", max_ctas)
        _, x0 = net.forward_stage(lrs, alphas, hm.stage_enc(0), (1, 64, size, size))
        _, y1 = net.forward_stage(lrs, alphas, hm.stage_enc(1), (1, 64, size, size))
        torch.cuda.synchronize()
        wb = w.to(torch.bfloat16).to(torch.float32)
        ref = _layer_ref(x0.cpu(), wb, params[key + ".bias"], params["encode.res_layers.0.block.1.weight"])
        err = (y1.cpu() - ref).abs()
        log(check="conv64_taps", max_ctas=max_ctas, size=size, taps=str(taps), max_err=float(err.max()),
            ref_max=float(ref.abs().max()), bad_frac=float((err > 2e-2 * ref.abs().max()).float().mean()))
        del net


def group_forward(case_names):
    import numpy as np
    import torch
    import highres_net_b200 as hb
    from highres_net_b200 import hrnet as hm
    from oracle import cases, hrnet_oracle
    dev = torch.device("cuda:0")
    params = hrnet_oracle.make_params(cases.WEIGHT_SEED)
    net = hb.HRNet(hrnet_oracle.DEFAULT_NETWORK_CONFIG).eval()
    net.load_state_dict(params)
    net = net.to(dev)
    golden = np.load(os.path.join(ROOT, "tests", "golden", "hrnet_forward.npz"))
    for name in case_names:
        lrs, alphas = cases.hrnet_inputs(name)
        b, l, s, _ = lrs.shape
        tr = {}
        hrnet_oracle.hrnet_forward(params, lrs, alphas, trace=tr)
        tl, ta = torch.from_numpy(lrs).to(dev), torch.from_numpy(alphas).to(dev)
        sr = net(tl, ta)
        torch.cuda.synchronize()
        log(check="forward_sr", case=name, max_err=float(np.abs(sr.cpu().numpy() - golden[name]).max()),
            ref_std=float(golden[name].std()))
        _, anchor = net.forward_stage(tl, ta, hm.stage_anchor(), (b, 1, s, s))
        log(check="anchor", case=name, max_err=float((anchor.cpu()[:, 0] - tr["anchor"]).abs().max()))
        n_enc = 2 * 2 + 1
        _, enc = net.forward_stage(tl, ta, hm.stage_enc(n_enc), (b * l, 64, s, s))
        ref = tr["encoded"].reshape(b * l, 64, s, s)
        log(check="encoder_out", case=name, max_err=float((enc.cpu() - ref).abs().max()), ref_max=float(ref.abs().max()))
        n, level = l, 0
        while n // 2 > 0:
            half = n // 2
            _, lv = net.forward_stage(tl, ta, hm.stage_fuse(level, 2), (b * half, 64, s, s))
            ref = tr["levels"][level].reshape(b * half, 64, s, s)
            log(check="fuse_level", case=name, level=level, max_err=float((lv.cpu() - ref).abs().max()),
                ref_max=float(ref.abs().max()))
            n, level = half, level + 1


def group_speed():
    """First look at throughput: C2-like batch slices, CUDA-event timing (not a bench number)."""
    import torch
    import highres_net_b200 as hb
    from oracle import hrnet_oracle
    dev = torch.device("cuda:0")
    net = hb.HRNet(hrnet_oracle.DEFAULT_NETWORK_CONFIG).eval()
    net.load_state_dict(hrnet_oracle.make_params(0))
    net = net.to(dev)
    for b, l, s in ((4, 16, 128), (32, 16, 128)):
        lrs = torch.rand(b, l, s, s, device=dev)
        al = torch.ones(b, l, device=dev)
        for _ in range(2):
            net(lrs, al)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters = 5
        e0.record()
        for _ in range(iters):
            net(lrs, al)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        fl = hrnet_oracle.flops_per_imageset(l, s, s) * b
        log(check="speed", B=b, L=l, S=s, ms=ms, imagesets_per_s=b / ms * 1e3, tflops=fl / ms / 1e9)


GROUPS = {
    "scoring": lambda: group_scoring(),
    "taps_s16": lambda: group_taps(0, 16),
    "taps_s128": lambda: group_taps(0, 128),
    "taps_s128_ctas5": lambda: group_taps(5, 128),
    "forward_small": lambda: group_forward(["b1_l1_s16", "b2_l4_s32", "b1_l5_s24", "b1_l6_s16", "b2_l9_s16", "b1_l16_s16"]),
    "forward_big": lambda: group_forward(["b1_l2_s136", "c1_b2_l4_s128"]),
    "speed": lambda: group_speed(),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--group")
    ap.add_argument("--groups", default=",".join(GROUPS))
    ap.add_argument("--timeout", type=int, default=240)
    args = ap.parse_args()
    if args.group:
        GROUPS[args.group]()
        return
    for g in args.groups.split(","):
        t0 = time.time()
        try:
            res = subprocess.run([sys.executable, os.path.abspath(__file__), "--group", g], capture_output=True,
                                 text=True, timeout=args.timeout)
            rc, out, err = res.returncode, res.stdout, res.stderr
        except subprocess.TimeoutExpired as e:
            rc, out, err = -999, (e.stdout or b"").decode() if isinstance(e.stdout, bytes) else (e.stdout or ""), "TIMEOUT"
        print(f"===== group {g}: rc={rc} ({time.time() - t0:.1f}s)")
        print(out, end="")
        if rc != 0:
            print("--- stderr tail ---")
            print("\n".join(err.splitlines()[-25:]))
        sys.stdout.flush()


if __name__ == "__main__":
    main()
