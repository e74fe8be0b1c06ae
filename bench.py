#!/usr/bin/env python
"""Headline benchmark: SR imagesets/sec of the HRNet inference hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--lean]

Workload (BASELINE.json configs[1]): HRNet inference, n_views=16, batch=32 per GPU,
128x128 LR -> 384x384 SR, synthetic inputs, random-init weights.  One "step" = one
pass of the hot path (HRNet.forward) over one batch.  N > 1 is launched by torchrun,
one rank per GPU; imagesets are sharded by batch (weak scaling: 32 imagesets per GPU).
The model has no data-path collective; the design's only collective -- the gather of
the SR images (and, in the scoring leg, of the (cPSNR, x, y) rows) over NCCL -- runs
every step on a side stream INSIDE the timed region when N > 1.  The max over ranks
of the device time is reported.

Prints ONE JSON line (see the build contract):
  value         whole-job imagesets/s with inputs resident in HBM
  e2e           the same metric through the public host-buffer API (pinned host lrs/alphas in, SR out, copies timed)
  roofline      dominant kernel class against the measured tensor peak; per-class times taken in the SAME sustained,
                power-capped regime as the timed loop, with the gaps between launches as their own class
  scoring       Lanczos shift and shifted-cPSNR search: GB/s against the HBM peak, the composite C4 path, CPU baselines
  configs       BASELINE.json configs[2] (per-rank shard) and configs[4]: throughput, fraction of roofline, SR error vs oracle
  gpu_baseline  the reference algorithm through PyTorch eager + cuDNN on the same GPU (TF32 and bf16 autocast)
  io            16-bit PNG imagesets on disk -> SR PNGs: native threaded decode / encode around the device path
  cpu_baseline  the oracle port (torch CPU fp32, all host threads) on a bounded sample
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "sr_imagesets_per_sec_16x128to384"
UNIT = "imagesets/s"
NCU_SUMMARY = os.path.join("profiles", "r02_ncu_full_summary.json")
LANCZOS_BYTES = 2 * 384 * 384 * 4          # read + write once per image
CPSNR_BYTES = 3 * 384 * 384 * 4            # sr, hr, map read once per imageset


def synthetic_batch(b, l, s, seed, device=None, pin=False):
    """lrs ~ U[0,1) (DataLoader.py:195-198 range), alphas = 1."""
    import torch
    g = torch.Generator().manual_seed(seed)
    lrs = torch.rand(b, l, s, s, generator=g, dtype=torch.float32)
    alphas = torch.ones(b, l, dtype=torch.float32)
    if device is not None:
        return lrs.to(device), alphas.to(device)
    if pin:
        return lrs.pin_memory(), alphas.pin_memory()
    return lrs, alphas


def _ncu_rows():
    """Per-launch rows of the committed `ncu --set full` capture of one forward step and whether the capture was taken
    from the library build that is running now (the summary records the source stamp of the build it profiled)."""
    path = os.path.join(ROOT, NCU_SUMMARY)
    if not os.path.exists(path):
        return None, None
    with open(path) as f:
        doc = json.load(f)
    rows = doc["launches"] if isinstance(doc, dict) else doc
    stamp_file = os.path.join(ROOT, "highres-net_b200", "csrc", "libhrn_b200.so.stamp")
    now = open(stamp_file).read().strip() if os.path.exists(stamp_file) else None
    fresh = isinstance(doc, dict) and now is not None and doc.get("lib_stamp") == now
    return rows, fresh


def _ncu_tag(kernel_class):
    return (kernel_class.replace("conv3x3_umma", "conv3x3_umma_kernel").replace("resblock64_umma", "resblock64_umma_kernel")
            .replace("fuse_wave", "fuse_wave_kernel").replace("enc_wave", "enc_wave_kernel").rstrip(">"))


def ncu_class_stats(kernel_class):
    """DRAM bytes (read + write) per launch and tensor-pipe activity of a kernel class from the committed capture."""
    rows, fresh = _ncu_rows()
    if rows is None:
        return None
    tag = _ncu_tag(kernel_class)
    sel = [r for r in rows if tag in r["name"]]
    if not sel:
        return None
    pipe = [float(r["tensor"]) for r in sel]
    return {"traffic": sum((float(r["rd"]) + float(r["wr"])) * 1e9 for r in sel) / len(sel),
            "tensor_pipe_active_pct": {"mean": sum(pipe) / len(pipe), "best_launch": max(pipe), "launches": len(sel)},
            "dram_bytes_per_forward": sum((float(r["rd"]) + float(r["wr"])) * 1e9 for r in rows),
            "source": NCU_SUMMARY + " (ncu --set full, one forward step, kernels serialised)",
            "capture_matches_running_build": bool(fresh)}


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"bf16_tflops": p.get("bf16_tflops_sustained", p.get("bf16_tflops")), "bf16_tflops_burst": p.get("bf16_tflops"),
                "hbm_gbs": p.get("hbm_gbs"), "source": "measured (MEASURED_PEAKS.json: sustained tensor peak for kernels "
                "timed inside the long step, burst HBM copy peak for the scoring kernels timed alone)"}
    return {"bf16_tflops": 1400.0, "bf16_tflops_burst": 1400.0, "hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (samples outside the
    [t0, t1] window handed to stop() are dropped)."""
    FIELDS = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(gpu_index), f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self, t0, t1):
        import datetime
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, mx, reasons, power = [], [], set(), []
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for line in out.strip().splitlines():
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 8:
                continue
            try:
                ts = datetime.datetime.strptime(parts[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                vals = (float(parts[1]), float(parts[2]), float(parts[3]))
            except ValueError:
                continue
            if ts < t0 or ts > t1:
                continue
            sm.append(vals[0])
            mx.append(vals[1])
            power.append(vals[2])
            for name, val in zip(names, parts[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_min_mhz": min(sm) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "power_w_median": statistics.median(power) if power else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ============================================================================ CPU legs (the oracle is the checker/baseline)
def cpu_baseline(l, s, seconds_budget=8.0, max_sets=24, threads=None):
    """Oracle port (torch CPU fp32) timed on the host cores; bounded sample of the same workload."""
    import torch
    from oracle import hrnet_oracle
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    params = hrnet_oracle.make_params(0)
    lrs, alphas = synthetic_batch(1, l, s, seed=99)
    t0 = time.perf_counter()
    hrnet_oracle.hrnet_forward(params, lrs, alphas)          # warm-up, also sizes the sample
    once = time.perf_counter() - t0
    n_sets = int(max(1, min(max_sets, seconds_budget // max(once, 1e-3))))      # x 2 repetitions: 10-20 s of CPU work
    lrs, alphas = synthetic_batch(n_sets, l, s, seed=100)
    best = float("inf")
    for _ in range(2):
        t0 = time.perf_counter()
        for i in range(n_sets):                                # batch 1 at a time: the reference val loop (train.py:284)
            hrnet_oracle.hrnet_forward(params, lrs[i:i + 1], alphas[i:i + 1])
        best = min(best, time.perf_counter() - t0)
    return {"value": n_sets / best, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{n_sets} imagesets of L={l} {s}x{s} fp32, oracle/hrnet_oracle.py on torch CPU "
                      f"(oneDNN), batch 1 per call, best of 2"}


def cpu_scoring_baselines(n_lanczos=8, n_cpsnr=4):
    """Oracle ports of lanczos_shift (lanczos.py:47-107) and shift_cPSNR (Evaluator.py:52-73) on the host: numpy, one
    thread, 384 x 384 images (BASELINE.md section 3 asks for these two next to HRNet)."""
    import numpy as np
    from oracle import scoring_oracle
    rng = np.random.RandomState(5)
    img = rng.rand(1, n_lanczos, 384, 384).astype(np.float32)
    shift = rng.uniform(-1, 1, size=(n_lanczos, 2)).astype(np.float32)
    scoring_oracle.lanczos_shift(img[:, :1], shift[:1], p=5)
    t0 = time.perf_counter()
    scoring_oracle.lanczos_shift(img, shift, p=5)
    t_l = time.perf_counter() - t0
    sr = rng.rand(n_cpsnr, 384, 384).astype(np.float32)
    hr = rng.rand(n_cpsnr, 384, 384).astype(np.float32)
    hm = (rng.rand(n_cpsnr, 384, 384) > 0.1).astype(np.float32)
    scoring_oracle.shift_cpsnr(sr[0], hr[0], hm[0])
    t0 = time.perf_counter()
    for i in range(n_cpsnr):                                   # one imageset per call, like predict.py:43
        scoring_oracle.shift_cpsnr(sr[i], hr[i], hm[i])
    t_c = time.perf_counter() - t0
    return ({"value": n_lanczos / t_l, "unit": "images/s", "cores": 1, "kind": "port",
             "sample": f"{n_lanczos} images of 384x384, oracle/scoring_oracle.py lanczos_shift (numpy)"},
            {"value": n_cpsnr / t_c, "unit": "imagesets/s", "cores": 1, "kind": "port",
             "sample": f"{n_cpsnr} imagesets of 384x384, oracle/scoring_oracle.py shift_cpsnr (numpy, 49 sites)"})


def run_reference(args, rank, world, out):
    """--impl reference: the reference's own CPU path (oracle port: the reference is pure PyTorch and
    /root/reference does not exist on the GPU box), all host threads, rank 0 only."""
    if rank != 0:
        return
    import torch
    from oracle import hrnet_oracle
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    params = hrnet_oracle.make_params(0)
    l, s = args.views, args.size
    sets_per_step = 2
    lrs, alphas = synthetic_batch(sets_per_step, l, s, seed=100)
    t0 = time.perf_counter()
    hrnet_oracle.hrnet_forward(params, lrs, alphas)          # one untimed warm-up step, also sizes the run
    once = time.perf_counter() - t0
    for _ in range(max(0, min(args.warmup, 3) - 1)):
        hrnet_oracle.hrnet_forward(params, lrs, alphas)
    steps = max(1, min(args.steps, int(150.0 / max(once, 1e-3))))   # all K steps unless that would take more than ~2.5 min
    t0 = time.perf_counter()
    for _ in range(steps):
        hrnet_oracle.hrnet_forward(params, lrs, alphas)
    dt = time.perf_counter() - t0
    value = sets_per_step * steps / dt
    sample = (f"{sets_per_step} imagesets of L={l} {s}x{s} per step, {steps} steps, oracle port of HRNet.forward on "
              f"torch CPU fp32 (oneDNN)")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": max(1, min(args.warmup, 3)), "ms_per_step": dt / steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"HRNet inference n_views={l} batch={args.batch}/GPU {s}x{s}->{3 * s}x{3 * s} "
                               f"(BASELINE.json configs[1]); reference arm times a bounded sample on host cores"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }), file=out, flush=True)


def _claim_stdout():
    """Keep stdout for the ONE JSON line: everything else that writes to fd 1 (e.g. NCCL's version banner, which comes
    from C code) is sent to stderr.  Returns a writer for the real stdout."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return real


# ============================================================================ GPU legs
class Bench:
    def __init__(self, args, rank, world, local_rank):
        import torch
        import torch.distributed as dist
        import highres_net_b200 as hb
        from oracle import hrnet_oracle  # parameter generator, flop model, checker and baseline legs only
        self.torch, self.dist, self.hb, self.oracle = torch, dist, hb, hrnet_oracle
        self.args, self.rank, self.world = args, rank, world
        torch.cuda.set_device(local_rank)
        self.dev = torch.device("cuda", local_rank)
        if world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=self.dev)
        self.params = hrnet_oracle.make_params(0)
        self.net = hb.HRNet(hrnet_oracle.DEFAULT_NETWORK_CONFIG).eval()
        self.net.load_state_dict(self.params)
        self.net = self.net.to(self.dev)
        if args.no_wave:
            self.net.debug_set(self.dev, "fuse_wave", 0)
        for kv in args.knob:
            name, value = kv.split("=")
            self.net.debug_set(self.dev, name, int(value))
        self.peaks = load_peaks()
        self.gather_stream = torch.cuda.Stream(self.dev) if world > 1 else None

    # ---- helpers
    def sync_all(self):
        self.torch.cuda.synchronize(self.dev)
        if self.world > 1:
            self.dist.barrier()
            self.torch.cuda.synchronize(self.dev)

    def max_over_ranks(self, values):
        if self.world == 1:
            return [float(v) for v in values]
        t = self.torch.tensor(values, dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(v) for v in t]

    def timed(self, fn, iters, warm=3, local=False):
        """ms per call of fn (CUDA events on the current stream); local=True: this rank alone (no barrier)."""
        torch = self.torch
        sync = (lambda: torch.cuda.synchronize(self.dev)) if local else self.sync_all
        for i in range(warm):
            fn(i)
        sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(iters):
            fn(i)
        if self.gather_stream is not None and not local:
            torch.cuda.current_stream(self.dev).wait_stream(self.gather_stream)
        e1.record()
        sync()
        return e0.elapsed_time(e1) / iters

    def ramp(self, inputs, seconds):
        """The board idles at 120 MHz and, once loaded, needs about a second to settle at its power-capped clock
        (sw_power_cap, ~900-1000 W).  Everything that follows a ramp is measured in that settled state."""
        t0 = time.time()
        while time.time() - t0 < seconds:
            for i in range(5):
                self.net(*inputs[i % len(inputs)])
            self.torch.cuda.synchronize(self.dev)

    def gather_sr(self, sr, buf):
        """The design's collective: all-gather of this rank's SR images on a side stream, so that it
        overlaps the next step's kernels."""
        torch = self.torch
        ev = torch.cuda.Event()
        ev.record()
        self.gather_stream.wait_event(ev)
        with torch.cuda.stream(self.gather_stream):
            self.dist.all_gather_into_tensor(buf, sr)
        sr.record_stream(self.gather_stream)

    # ---- the headline loops
    def headline(self):
        torch, args, net, dev, world = self.torch, self.args, self.net, self.dev, self.world
        b, l, s = args.batch, args.views, args.size
        n_rot = max(2, int(140e6 // (b * l * s * s * 4)) + 1)       # rotate over > 126 MB of distinct inputs
        self.n_rot = n_rot
        self.dev_inputs = [synthetic_batch(b, l, s, seed=1 + self.rank * 1000 + i, device=dev) for i in range(n_rot)]
        host_inputs = [synthetic_batch(b, l, s, seed=1 + self.rank * 1000 + i, pin=True) for i in range(min(n_rot, 3))]
        host_outs = [torch.empty((b, 1, 3 * s, 3 * s), dtype=torch.float32).pin_memory() for _ in range(2)]
        gbuf = torch.empty((world * b, 1, 3 * s, 3 * s), dtype=torch.float32, device=dev) if world > 1 else None
        net.reserve(dev, b, l, s, s)
        self.ramp(self.dev_inputs, 2.0)

        sampler = ClockSampler(dev.index) if self.rank == 0 else None
        for i in range(args.warmup):
            sr = net(*self.dev_inputs[i % n_rot])
            if world > 1:
                self.gather_sr(sr, gbuf)
        self.sync_all()
        launches0 = self.hb.kernel_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_wall0 = time.time()
        e0.record()
        for i in range(args.steps):
            sr = net(*self.dev_inputs[i % n_rot])
            if world > 1:
                self.gather_sr(sr, gbuf)
        if world > 1:
            torch.cuda.current_stream(dev).wait_stream(self.gather_stream)
        e1.record()
        self.sync_all()
        t_wall1 = time.time()
        launches = self.hb.kernel_launch_count() - launches0
        ms_total = e0.elapsed_time(e1)
        clocks = sampler.stop(t_wall0, t_wall1) if sampler is not None else None

        # ---- end to end through the host-buffer API: the validation-loop pattern (train.py:199-208) through
        # forward_host_submit / forward_host_wait: every step copies its own inputs from pinned host memory and its SR back
        # into pinned host memory; two steps are in flight, so the copies of step n +- 1 overlap the kernels of step n.
        def e2e_loop(n):
            pending = None
            for i in range(n):
                nxt = net.forward_host_submit(*host_inputs[i % len(host_inputs)], out_host=host_outs[i % 2], device=dev)
                if pending is not None:
                    net.forward_host_wait(pending)
                pending = nxt
            net.forward_host_wait(pending)

        e2e_loop(3)
        self.sync_all()
        e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e2.record()
        e2e_loop(args.steps)
        e3.record()
        self.sync_all()
        ms_e2e = e2.elapsed_time(e3)
        e4, e5 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e4.record()
        for i in range(args.steps):                     # one call at a time: copy in, kernels, copy out, nothing overlapped
            net.forward_host(*host_inputs[i % len(host_inputs)], out_host=host_outs[0], device=dev)
        e5.record()
        self.sync_all()
        ms_e2e_sync = e4.elapsed_time(e5)
        self.local_ms_total = ms_total                            # this rank's own clock: what its per-launch profile is compared with
        ms_total, ms_e2e, ms_e2e_sync = self.max_over_ranks([ms_total, ms_e2e, ms_e2e_sync])
        if world > 1:
            lt = torch.tensor([launches], dtype=torch.int64, device=dev)
            self.dist.all_reduce(lt)
            launches = int(lt[0])
        return ms_total, ms_e2e, ms_e2e_sync, launches, clocks

    def sustained_profile(self):
        """Per-launch CUDA events (hrn_profile_begin/_end) in the SAME regime as the timed loop: a fresh ramp to the
        power-capped clock, then as many profiled steps as the timed loop had.  The `forward_span` class brackets each
        whole forward, so forward_span - sum(kernel classes) = time between launches (`gap`)."""
        self.ramp(self.dev_inputs, 1.0)
        steps = self.args.steps
        self.net.profile_begin(self.dev)
        for i in range(steps):
            self.net(*self.dev_inputs[i % self.n_rot])
        prof = self.net.profile_end(self.dev)
        span = prof.pop("forward_span")
        kernels_ms = sum(v["ms"] for v in prof.values())
        prof["gap"] = {"ms": max(0.0, span["ms"] - kernels_ms), "flops": 0.0, "launches": 0}
        return prof, span["ms"] / steps, steps

    # ---- other BASELINE configs
    def config_leg(self, b, l, s, iters, check_row, gate):
        torch, net, dev = self.torch, self.net, self.dev
        g = torch.Generator().manual_seed(4242 + l + s)
        lrs = torch.rand(b, l, s, s, generator=g)
        alphas = torch.ones(b, l)
        tl, ta = lrs.to(dev), alphas.to(dev)
        sr = net(tl, ta)
        err = None
        if self.rank == 0:
            ref = self.oracle.hrnet_forward(self.params, lrs[check_row:check_row + 1].numpy(),
                                            alphas[check_row:check_row + 1].numpy()).numpy()
            err = float(abs(sr[check_row:check_row + 1].cpu().numpy() - ref).max())
        ms = self.timed(lambda i: net(tl, ta), iters)
        ms, = self.max_over_ranks([ms])
        del tl, ta, sr
        torch.cuda.empty_cache()
        value = b * self.world / ms * 1e3
        tf = self.oracle.flops_per_imageset(l, s, s) * b / ms / 1e9
        return {"workload": f"B{b}/GPU L{l} {s}x{s}->{3 * s}x{3 * s}", "value": value, "unit": UNIT, "ms_per_step": ms,
                "model_tflops_per_gpu": tf, "frac_of_roofline": tf / self.peaks["bf16_tflops"],
                "frac_of_burst_roofline": tf / self.peaks["bf16_tflops_burst"],
                "sr_max_err": err, "sr_gate": gate, "parity_ok": (err is not None and err <= gate),
                "sr_check": f"imageset {check_row} of the batch against oracle/hrnet_oracle.py (fp32 CPU), rank 0",
                "timing": f"{iters} steps after 3 warm-up steps, CUDA events, max over ranks, same inputs every step "
                          f"(activations alone are {5 * b * l * s * s * 128 / 1e9:.1f} GB per step >> L2)"}

    # ---- scoring kernels and the composite C4 path
    def scoring_leg(self):
        torch, hb, dev, world, net = self.torch, self.hb, self.dev, self.world, self.net
        hbm = self.peaks["hbm_gbs"]
        out = {}
        n = 512
        big = torch.rand(1, n, 384, 384, device=dev)
        sh = torch.rand(n, 2, device=dev) * 2 - 1
        ms = self.timed(lambda i: hb.lanczos_shift(big, sh, p=5), 200, warm=20)
        gbps = n * LANCZOS_BYTES / ms / 1e6
        out["lanczos"] = {"GBps": gbps, "frac": gbps / hbm, "ms": ms, "images": n, "bytes_per_image": LANCZOS_BYTES,
                          "traffic": self._scoring_traffic("lanczos7_tma_kernel", n)}
        srb, hrb = big[0], torch.rand(n, 384, 384, device=dev)
        hmb = (torch.rand(n, 384, 384, device=dev) > 0.1).float()
        ms = self.timed(lambda i: hb.shift_cPSNR_argmax(srb, hrb, hmb), 40, warm=5)
        gbps = n * CPSNR_BYTES / ms / 1e6
        out["cpsnr"] = {"GBps_algorithmic": gbps, "frac": gbps / hbm, "ms": ms, "imagesets": n,
                        "bytes_per_imageset": CPSNR_BYTES, "imagesets_per_s": n / ms * 1e3,
                        "traffic": self._scoring_traffic("cpsnr_onepass_kernel", n),
                        "note": "one pass over sr, hr and the map: n, sum(m d), sum(m d^2) for 49 shifts from centred fp32 "
                                "partial sums, 4 fp32 instructions per (shift, pixel) = ~200 per pixel: bound by fp32 issue, not "
                                "by HBM (DESIGN.md); frac is against the HBM copy peak as SURVEY 8d asks"}
        del big, hrb, hmb, srb
        # C4: BASELINE.json configs[3] -- forward + lanczos_shift + clip + shift_cPSNR on 32 x 16-view imagesets per
        # GPU; at N > 1 the (cPSNR, x, y) rows of every rank are all-gathered every step and the SR images on the side stream
        b, l, s = self.args.batch, 16, 128
        tl, ta = synthetic_batch(b, l, s, seed=77 + self.rank, device=dev)
        g = torch.Generator().manual_seed(78)
        shift = (torch.rand(b, 2, generator=g) * 2 - 1).to(dev)
        hm = (torch.rand(b, 3 * s, 3 * s, generator=g) > 0.1).float().to(dev)
        sr0 = net(tl, ta)[:, 0]
        # hr = clip(roll(clip(sr)) + 0.02 + N(0, sigma^2)), sigma = 0.01 (about 40 dB at the known shift)
        noise = (0.01 * torch.randn(b, 3 * s, 3 * s, generator=g)).to(dev)
        hr = (torch.roll(hb.lanczos_shift(sr0[None], shift, p=5)[0].clamp(0, 1), (1, -2), (1, 2)) + 0.02 + noise).clamp(0, 1)
        gbuf = torch.empty((world * b, 1, 3 * s, 3 * s), dtype=torch.float32, device=dev) if world > 1 else None
        rows = torch.empty((world * b, 3), dtype=torch.float32, device=dev) if world > 1 else None

        def c4(i):
            sr = net(tl, ta)
            moved = hb.lanczos_shift(sr[:, 0][None], shift, p=5)[0]
            best, xy, _ = hb.shift_cPSNR_argmax(moved, hr, hm, clip_sr=True)
            if world > 1:
                packed = torch.cat([best[:, None], xy.float()], 1)
                self.dist.all_gather_into_tensor(rows, packed)
                self.gather_sr(sr, gbuf)
            return best, xy

        best, xy = c4(0)
        ms_c4 = self.timed(c4, 20)
        ms_fw = self.timed(lambda i: net(tl, ta), 20)
        ms_c4, ms_fw = self.max_over_ranks([ms_c4, ms_fw])
        out["c4_imagesets_per_s"] = b * world / ms_c4 * 1e3
        out["c4"] = {"ms_per_step": ms_c4, "forward_only_ms": ms_fw, "scoring_share": max(0.0, 1.0 - ms_fw / ms_c4),
                     "best_shift_is_the_known_roll": bool((xy[:, 0] == 4).all() and (xy[:, 1] == 1).all()),
                     "collective": ("all_gather of (cPSNR, x, y) rows every step + SR images on a side stream (NCCL)"
                                    if world > 1 else "none at N=1")}
        return out

    def _scoring_traffic(self, kernel, units):
        path = os.path.join(ROOT, "profiles", "r02_ncu_scoring_summary.json")
        if not os.path.exists(path):
            return None
        with open(path) as f:
            doc = json.load(f)
        rows = [r for r in (doc["launches"] if isinstance(doc, dict) else doc) if kernel in r["name"]]
        if not rows:
            return None
        per_call = sum((float(r["rd"]) + float(r["wr"])) * 1e9 for r in rows) / max(1, doc.get("calls", 1) if isinstance(doc, dict) else 1)
        return {"dram_bytes_per_call": per_call, "dram_bytes_per_unit": per_call / (doc.get("units", units) if isinstance(doc, dict) else units),
                "source": "profiles/r02_ncu_scoring_summary.json"}

    # ---- disk -> SR: the file formats on both ends around the device path
    def io_leg(self, n_sets=64, l=16, s=128):
        import shutil
        import tempfile
        import numpy as np
        from highres_net_b200 import imageset_io as io
        from highres_net_b200.predict import collate_device, img_as_uint_u16
        torch, dev, net = self.torch, self.dev, self.net
        root = tempfile.mkdtemp(prefix="hrn_io_")
        try:
            rng = np.random.RandomState(0)
            dirs = []
            base = (rng.rand(l, s, s) * 0.5 * 65535).astype(np.uint16)
            qm = np.full((l, s, s), 255, np.uint16)
            for i in range(n_sets):                     # synthetic Proba-V-shaped imagesets: L views of s x s, 16-bit PNG
                d = os.path.join(root, f"imgset{i:04d}")
                os.makedirs(d)
                views = base + rng.randint(0, 2000, size=(l, 1, 1)).astype(np.uint16)
                io.write_png_u16([os.path.join(d, f"LR{v:03d}.png") for v in range(l)], views)
                io.write_png_u16([os.path.join(d, f"QM{v:03d}.png") for v in range(l)], qm)
                io.write_png_u16([os.path.join(d, "SM.png")], np.full((1, 3 * s, 3 * s), 255, np.uint16))
                dirs.append(d)
            io.save_clearance_scores(dirs)
            ds = io.ImagesetDataset(dirs, {"create_patches": False, "patch_size": 64}, raw16=True)
            lr_bytes = sum(os.path.getsize(os.path.join(d, f)) for d in dirs for f in os.listdir(d) if f.startswith("LR"))
            t0 = time.perf_counter()
            sets = [ds[i] for i in range(n_sets)]       # one imageset per call: PNG decode on the native thread pool, clearance order
            t_read1 = time.perf_counter() - t0
            t0 = time.perf_counter()
            sets = [im for b0 in range(0, n_sets, 32) for im in ds[b0:b0 + 32]]     # 32 imagesets (512 files) per native call
            t_read = time.perf_counter() - t0
            out_dir = os.path.join(root, "sr")
            os.makedirs(out_dir)
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            for b0 in range(0, n_sets, 32):
                group = ds[b0:min(n_sets, b0 + 32)]
                lrs, alphas, _, _, names = collate_device(group, l, dev)
                sr16 = img_as_uint_u16(net(lrs, alphas)[:, 0]).cpu()
                io.write_png_u16([os.path.join(out_dir, n + ".png") for n in names], sr16)
            t_all = time.perf_counter() - t0
            del sets
            return {"imagesets": n_sets, "views_per_imageset": l, "lr_png_bytes": lr_bytes, "host_threads": os.cpu_count(),
                    "png_decode_views_per_s": n_sets * l / t_read, "png_decode_MBps_decoded": n_sets * l * s * s * 2 / t_read / 1e6,
                    "png_decode_views_per_s_one_imageset_per_call": n_sets * l / t_read1,
                    "disk_to_sr_png_imagesets_per_s": n_sets / t_all,
                    "what": "16-bit PNG views on local disk -> native threaded decode -> pinned uint16 -> hrn_collate -> HRNet -> "
                            "img_as_uint on the device -> native PNG encode of the 384x384 SR (DataLoader.py:73-148, predict.py:161-194)"}
        finally:
            shutil.rmtree(root, ignore_errors=True)

    # ---- the reference algorithm through PyTorch eager + cuDNN on this GPU (SURVEY.md sections 2a / 8d)
    def gpu_baseline_leg(self):
        torch, dev = self.torch, self.dev
        b, l, s = self.args.batch, self.args.views, self.args.size
        params = {k: v.to(dev) for k, v in self.params.items()}
        lrs, al = synthetic_batch(b, l, s, seed=5, device=dev)
        prev = torch.backends.cudnn.benchmark
        torch.backends.cudnn.benchmark = True
        res = {"what": "oracle/hrnet_oracle.py (the reference's op sequence: torch.nn.functional convs -> cuDNN) on the same "
                       "GPU; fp32 storage with TF32 convs is what the unmodified reference does on a GPU (predict.py:35-39)"}

        def run(fn, key):
            try:
                ms = self.timed(lambda i: fn(), 5, warm=2, local=True)
                res[key] = {"imagesets_per_s": b / ms * 1e3, "ms_per_step": ms}
            except Exception as e:          # noqa: BLE001  (an OOM or a missing cuDNN engine must not kill the bench line)
                res[key] = {"error": str(e)[:200]}

        with torch.no_grad():
            run(lambda: self.oracle.hrnet_forward(params, lrs, al), "torch_eager_tf32")

            def autocast_run():
                with torch.autocast("cuda", dtype=torch.bfloat16):
                    return self.oracle.hrnet_forward(params, lrs, al)
            run(autocast_run, "torch_bf16_autocast")
        torch.backends.cudnn.benchmark = prev
        del params, lrs, al
        torch.cuda.empty_cache()
        return res


def main():
    out = _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=32, help="imagesets per GPU per step")
    ap.add_argument("--views", type=int, default=16)
    ap.add_argument("--size", type=int, default=128)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--lean", action="store_true", help="headline loops and roofline only (A/B runs, ncu captures)")
    ap.add_argument("--no-wave", action="store_true", help="fusion levels as three launches each instead of one wavefront launch (A/B)")
    ap.add_argument("--knob", action="append", default=[], metavar="NAME=VALUE", help="hrn_debug_set knob for A/B runs (repeatable)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world, out)
        return

    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 path has no CPU fallback (use --impl reference for the CPU arm)")
    bench = Bench(args, rank, world, local_rank)
    b, l, s = args.batch, args.views, args.size
    flops_set = bench.oracle.flops_per_imageset(l, s, s)

    ms_total, ms_e2e, ms_e2e_sync, launches, clocks = bench.headline()
    prof, prof_step_ms, prof_steps = bench.sustained_profile()
    extras = {}
    if not args.lean:
        extras["configs"] = {
            "c3_shard": bench.config_leg(32, 32, 128, 8, check_row=7, gate=1e-2),
            "c5": bench.config_leg(8, 8, 512, 8, check_row=3, gate=1e-2),
        }
        extras["scoring"] = bench.scoring_leg()
        if rank == 0:
            extras["gpu_baseline"] = bench.gpu_baseline_leg()
            if world == 1:
                extras["io"] = bench.io_leg()
        bench.sync_all()

    if rank == 0:
        peaks = bench.peaks
        total_sets = b * world * args.steps
        ms_step = ms_total / args.steps
        value = total_sets / ms_total * 1e3
        e2e_value = total_sets / ms_e2e * 1e3
        tensor_classes = ("conv3x3_umma<64>", "conv3x3_umma<128>", "resblock64_umma", "fuse_wave", "enc_wave")
        dom = max(tensor_classes, key=lambda k: prof[k]["ms"])
        ach = prof[dom]["flops"] / max(prof[dom]["ms"], 1e-9) / 1e9       # TFLOP/s
        ncu = ncu_class_stats(dom) or {}
        attributed = sum(v["ms"] for k, v in prof.items() if k != "gap") / prof_steps
        roofline = {
            "bound": "tensor", "kernel": dom, "achieved": ach, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
            "frac": ach / peaks["bf16_tflops"], "frac_of_burst_peak": ach / peaks["bf16_tflops_burst"],
            "peak_burst": peaks["bf16_tflops_burst"], "peak_source": peaks["source"],
            "traffic": ncu.get("traffic"), "traffic_capture_matches_running_build": ncu.get("capture_matches_running_build"),
            "tensor_pipe_active_pct": ncu.get("tensor_pipe_active_pct"),
            "dram_bytes_per_forward": ncu.get("dram_bytes_per_forward"),
            "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum per launch, mean over the launches of this class in one "
                            f"forward step ({NCU_SUMMARY}); stale when traffic_capture_matches_running_build is false",
            "avg_launch_ms": prof[dom]["ms"] / max(prof[dom]["launches"], 1),
            "share_of_step": prof[dom]["ms"] / prof_steps / max(prof_step_ms, 1e-9),
            "whole_step": {"model_tflops": flops_set * b / ms_step / 1e9,
                           "frac": flops_set * b / ms_step / 1e9 / peaks["bf16_tflops"],
                           "frac_of_burst_peak": flops_set * b / ms_step / 1e9 / peaks["bf16_tflops_burst"],
                           "target_frac": 0.90},
            "attribution": {
                "regime": f"{prof_steps} profiled steps right after a 1 s ramp to the power-capped clock (the regime of the timed "
                          f"loop); per-launch CUDA events on the launching stream (events between launches switch off the "
                          f"programmatic-dependent-launch overlap, so a profiled step is never shorter than a timed one)",
                "profiled_step_ms": prof_step_ms, "timed_step_ms": ms_step,
                "timed_step_ms_rank0": bench.local_ms_total / args.steps,
                "kernel_classes_ms": attributed, "gap_ms": prof["gap"]["ms"] / prof_steps,
                # the profile is rank 0's, so it is held against rank 0's own timed step (ms_per_step is the max over ranks)
                "attributed_frac_of_timed_step": (attributed + prof["gap"]["ms"] / prof_steps) / (bench.local_ms_total / args.steps),
                "ok": (attributed + prof["gap"]["ms"] / prof_steps) >= 0.97 * (bench.local_ms_total / args.steps)},
            "per_class": {k: {"ms_per_step": v["ms"] / prof_steps, "launches_per_step": v["launches"] / prof_steps,
                              "share_of_profiled_step": v["ms"] / prof_steps / max(prof_step_ms, 1e-9),
                              "tflops": (v["flops"] / max(v["ms"], 1e-9) / 1e9) if v["flops"] else None}
                          for k, v in prof.items()},
        }
        if not roofline["attribution"]["ok"]:
            print("bench.py: WARNING: per-class times + gaps explain less than 97 % of the timed step", file=sys.stderr)
        # Tensor-pipe utilisation at the clock the timed loop actually ran at: one SM issues 8192 dense bf16 flop per
        # cycle (tcgen05.mma M128 N192 K16 = 96.1 cycles, tools/umma_probe.cu), 148 SMs.
        if clocks and clocks.get("sm_mhz"):
            per_ghz = 148 * 8192 * 1e-3                         # TFLOP/s per GHz of SM clock at 100 % pipe activity
            tf_clock = per_ghz * clocks["sm_mhz"] * 1e-3
            roofline["tensor_pipe_util_at_sampled_clock"] = {
                "sm_mhz": clocks["sm_mhz"], "pipe_peak_tflops_at_clock": tf_clock,
                "dominant_kernel": ach / tf_clock,
                "whole_step": flops_set * b / ms_step / 1e9 / tf_clock,
                "note": "achieved TFLOP/s / (148 SMs x 8192 flop/cycle x median SM clock of rank 0 during the timed loop)"}
        which = {(32, 16, 128): "BASELINE.json configs[1]", (32, 32, 128): "BASELINE.json configs[2], per-GPU shard",
                 (2, 4, 128): "BASELINE.json configs[0] shape", (8, 8, 512): "BASELINE.json configs[4] shape"}.get(
                     (b, l, s), "non-default shape")
        sr_bytes = b * 9 * s * s * 4
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"HRNet inference n_views={l} batch={b}/GPU {s}x{s}->{3 * s}x{3 * s} "
                                   f"({which}), random-init weights, fp32 in/out, bf16 activations "
                                   f"with fp32 accumulation",
                       "l2": f"inputs rotate over {bench.n_rot} distinct batches ({bench.n_rot * b * l * s * s * 4 / 1e6:.0f} MB "
                             f"> 126 MB L2); activations stream through HBM every step",
                       "sharding": "batch (independent imagesets per rank); no collective inside the model",
                       "fusion_schedule": ("three launches per level (--no-wave)" if args.no_wave else
                                           "one wavefront launch per fusion level (fuse_wave_umma.cu)"),
                       "knobs": args.knob,
                       "collective": (f"every step: NCCL all_gather of the SR images ({sr_bytes} B per rank, {world * sr_bytes} B "
                                      f"gathered per rank) on a side stream, inside the timed region" if world > 1 else
                                      "none at N=1 (the SR/score gather runs when N > 1)"),
                       "power": "this workload runs at the board power cap (see clocks: sw_power_cap, SM clock well below "
                                "max); the device-resident loop holds the cap continuously, while the copy phases of the e2e "
                                "loop let the clocks recover, so e2e can match or exceed value on the same box"},
            "model_tflops": flops_set * total_sets / ms_total / 1e9,
            "roofline": roofline,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": (b * l * s * s + b * l) * 4,
                    "d2h_bytes_per_step": sr_bytes, "ms_per_step": ms_e2e / args.steps,
                    "api": "HRNet.forward_host_submit / forward_host_wait (hrn_forward_host_submit): pinned host lrs/alphas "
                           "in, pinned host sr out, every step; two steps in flight",
                    "one_call_at_a_time": {"value": total_sets / ms_e2e_sync * 1e3, "ms_per_step": ms_e2e_sync / args.steps,
                                           "api": "HRNet.forward_host (hrn_forward_host), synchronous"}},
            "gpu_launches": launches,
            "clocks": clocks,
        }
        line.update(extras)
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(l, s)
            if "scoring" in line:
                lz, cp = cpu_scoring_baselines()
                line["scoring"]["cpu_baseline_lanczos"] = lz
                line["scoring"]["cpu_baseline_shift_cpsnr"] = cp
        else:
            line["cpu_baseline"] = None
        print(json.dumps(line), file=out, flush=True)
    if world > 1:
        bench.dist.destroy_process_group()


if __name__ == "__main__":
    main()
