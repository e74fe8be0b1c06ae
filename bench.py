#!/usr/bin/env python
"""Headline benchmark: SR imagesets/sec of the HRNet inference hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

Workload (BASELINE.json configs[1]): HRNet inference, n_views=16, batch=32 per GPU,
128x128 LR -> 384x384 SR, synthetic inputs, random-init weights.  One "step" = one
pass of the hot path (HRNet.forward) over one batch.  N > 1 is launched by torchrun,
one rank per GPU; imagesets are sharded by batch (weak scaling: 32 imagesets per GPU),
no data-path collective, the max over ranks of the device time is reported.

Prints ONE JSON line (see the build contract): value = whole-job imagesets/s with
inputs resident in HBM; e2e = the same metric through the public host-buffer API
(pinned host lrs/alphas in, SR out, copies inside the timed region); roofline = the
dominant kernel class against the measured tensor peak; cpu_baseline = the oracle
port (torch CPU fp32, all host threads) on a bounded sample.
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


def _ncu_tag(kernel_class):
    # conv3x3_umma<128> -> "conv3x3_umma_kernel<128" (the kernel has further template arguments: <CIN, POOL, MCAST>)
    return kernel_class.replace("conv3x3_umma", "conv3x3_umma_kernel").replace("resblock64_umma", "resblock64_umma_kernel").rstrip(">")


def ncu_traffic_per_launch(kernel_class):
    """DRAM bytes (read + write) per launch of a kernel class, averaged over the launches of one forward step, from the
    committed `ncu --set full` capture (profiles/r01_ncu_full_summary.json); None if the capture is not there."""
    path = os.path.join(ROOT, "profiles", "r01_ncu_full_summary.json")
    if not os.path.exists(path):
        return None
    tag = _ncu_tag(kernel_class)
    with open(path) as f:
        rows = [r for r in json.load(f) if tag in r["name"]]
    if not rows:
        return None
    return sum((float(r["rd"]) + float(r["wr"])) * 1e9 for r in rows) / len(rows)


def ncu_tensor_pipe_active(kernel_class):
    """sm__pipe_tensor_cycles_active (% of elapsed) of a kernel class from the same committed capture: mean over its
    launches in one forward step and the best single launch; None if the capture is not there."""
    path = os.path.join(ROOT, "profiles", "r01_ncu_full_summary.json")
    if not os.path.exists(path):
        return None
    tag = _ncu_tag(kernel_class)
    with open(path) as f:
        vals = [float(r["tensor"]) for r in json.load(f) if tag in r["name"]]
    if not vals:
        return None
    return {"mean": sum(vals) / len(vals), "best_launch": max(vals), "launches": len(vals),
            "source": "profiles/r01_ncu_full_summary.json (ncu --set full, one forward step, kernels serialised)"}


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"bf16_tflops": p.get("bf16_tflops_sustained", p.get("bf16_tflops")), "hbm_gbs": p.get("hbm_gbs"),
                "source": "measured (MEASURED_PEAKS.json, sustained)"}
    return {"bf16_tflops": 1400.0, "hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}


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
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world, out)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 path has no CPU fallback (use --impl reference for the CPU arm)")
    import highres_net_b200 as hb
    from oracle import hrnet_oracle  # parameter generator + flop model only; the oracle is timed in cpu_baseline alone

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    b, l, s = args.batch, args.views, args.size
    net = hb.HRNet(hrnet_oracle.DEFAULT_NETWORK_CONFIG).eval()
    net.load_state_dict(hrnet_oracle.make_params(0))
    net = net.to(dev)

    # inputs: rotate over enough distinct batches to exceed the 126 MB L2
    n_rot = max(2, int(140e6 // (b * l * s * s * 4)) + 1)
    dev_inputs = [synthetic_batch(b, l, s, seed=1 + rank * 1000 + i, device=dev) for i in range(n_rot)]
    host_inputs = [synthetic_batch(b, l, s, seed=1 + rank * 1000 + i, pin=True) for i in range(min(n_rot, 3))]
    host_out = torch.empty((b, 1, 3 * s, 3 * s), dtype=torch.float32).pin_memory()

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    # ---------------- clock / power ramp (untimed) ----------------
    # The board idles at 120 MHz and, once loaded, needs about a second to settle at its power-capped clock
    # (sw_power_cap, ~900-1000 W).  Everything below is measured in that settled state.
    t_ramp = time.time()
    while time.time() - t_ramp < 2.0:
        for i in range(5):
            net(*dev_inputs[i % n_rot])
        torch.cuda.synchronize(dev)

    # ---------------- device-resident throughput ----------------
    sampler = ClockSampler(local_rank) if rank == 0 else None
    for i in range(args.warmup):
        net(*dev_inputs[i % n_rot])
    sync_all()
    launches0 = hb.kernel_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.time()
    e0.record()
    for i in range(args.steps):
        sr = net(*dev_inputs[i % n_rot])
    e1.record()
    sync_all()
    t_wall1 = time.time()
    launches = hb.kernel_launch_count() - launches0
    ms_total = e0.elapsed_time(e1)
    clocks = sampler.stop(t_wall0, t_wall1) if sampler is not None else None

    # ---------------- end to end through the host-buffer API ----------------
    # The validation-loop pattern (train.py:199-208) through forward_host_submit / forward_host_wait: every step copies
    # its own inputs from pinned host memory and its SR back into pinned host memory; two steps are in flight, so the
    # copies of step n +- 1 overlap the kernels of step n.  The region ends when the last step's SR is on the host.
    host_outs = [host_out, torch.empty_like(host_out).pin_memory()]

    def e2e_loop(n):
        pending = None
        for i in range(n):
            nxt = net.forward_host_submit(*host_inputs[i % len(host_inputs)], out_host=host_outs[i % 2], device=dev)
            if pending is not None:
                net.forward_host_wait(pending)
            pending = nxt
        net.forward_host_wait(pending)

    e2e_loop(3)
    sync_all()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    e2e_loop(args.steps)
    e3.record()
    sync_all()
    ms_e2e = e2.elapsed_time(e3)
    # the same loop one call at a time (forward_host: copy in, kernels, copy out, nothing overlapped)
    e4, e5 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e4.record()
    for i in range(args.steps):
        net.forward_host(*host_inputs[i % len(host_inputs)], out_host=host_out, device=dev)
    e5.record()
    sync_all()
    ms_e2e_sync = e4.elapsed_time(e5)

    # ---------------- per-kernel-class timing for the roofline ----------------
    net.profile_begin(dev)
    prof_steps = min(args.steps, 5)
    for i in range(prof_steps):
        net(*dev_inputs[i % n_rot])
    prof = net.profile_end(dev)

    if world > 1:
        t = torch.tensor([ms_total, ms_e2e, ms_e2e_sync], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total, ms_e2e, ms_e2e_sync = float(t[0]), float(t[1]), float(t[2])
        lt = torch.tensor([launches], dtype=torch.int64, device=dev)
        dist.all_reduce(lt)
        launches = int(lt[0])

    if rank == 0:
        peaks = load_peaks()
        total_sets = b * world * args.steps
        value = total_sets / ms_total * 1e3
        e2e_value = total_sets / ms_e2e * 1e3
        dom = max(("conv3x3_umma<64>", "conv3x3_umma<128>", "resblock64_umma"), key=lambda k: prof[k]["ms"])
        ach = prof[dom]["flops"] / max(prof[dom]["ms"], 1e-9) / 1e9       # TFLOP/s
        step_ms = sum(v["ms"] for v in prof.values()) / prof_steps
        roofline = {
            "bound": "tensor", "kernel": dom, "achieved": ach, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
            "frac": ach / peaks["bf16_tflops"], "peak_source": peaks["source"], "traffic": ncu_traffic_per_launch(dom),
            "tensor_pipe_active_pct": ncu_tensor_pipe_active(dom),
            "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum per launch, mean over the launches of this class in one "
                            "forward step (profiles/r01_ncu_full_summary.json); equals the activation bytes read + written once",
            "avg_launch_ms": prof[dom]["ms"] / max(prof[dom]["launches"], 1),
            "share_of_step": prof[dom]["ms"] / prof_steps / max(step_ms, 1e-9),
            "per_class": {k: {"ms_per_step": v["ms"] / prof_steps, "launches_per_step": v["launches"] / prof_steps,
                              "tflops": (v["flops"] / max(v["ms"], 1e-9) / 1e9) if v["flops"] else None}
                          for k, v in prof.items()},
        }
        # Tensor-pipe utilisation at the clock the timed loop actually ran at: one SM issues 8192 dense bf16 flop per
        # cycle (tcgen05.mma M128 N192 K16 = 96.1 cycles, tools/umma_probe.cu), 148 SMs.  The ncu figure above is taken
        # with the kernels serialised at ~1.9 GHz (no power cap), where memory latency weighs more.
        if clocks and clocks.get("sm_mhz"):
            per_ghz = 148 * 8192 * 1e-3                         # TFLOP/s per GHz of SM clock at 100 % pipe activity
            tf_clock = per_ghz * clocks["sm_mhz"] * 1e-3
            roofline["tensor_pipe_util_at_sampled_clock"] = {
                "sm_mhz": clocks["sm_mhz"], "pipe_peak_tflops_at_clock": tf_clock,
                "dominant_kernel": ach / tf_clock,
                "whole_step": hrnet_oracle.flops_per_imageset(l, s, s) * b * world * args.steps / ms_total / 1e9 / world / tf_clock,
                "note": "achieved TFLOP/s / (148 SMs x 8192 flop/cycle x median SM clock of rank 0 during the timed loop); "
                        "dominant_kernel uses the per-class time of the profile loop that follows the timed loop"}
        flops_step = hrnet_oracle.flops_per_imageset(l, s, s) * b
        which = {(32, 16, 128): "BASELINE.json configs[1]", (32, 32, 128): "BASELINE.json configs[2], per-GPU shard",
                 (2, 4, 128): "BASELINE.json configs[0] shape", (8, 8, 512): "BASELINE.json configs[4] shape"}.get(
                     (b, l, s), "non-default shape")
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"HRNet inference n_views={l} batch={b}/GPU {s}x{s}->{3 * s}x{3 * s} "
                                   f"({which}), random-init weights, fp32 in/out, bf16 activations "
                                   f"with fp32 accumulation",
                       "l2": f"inputs rotate over {n_rot} distinct batches ({n_rot * b * l * s * s * 4 / 1e6:.0f} MB "
                             f"> 126 MB L2); activations stream through HBM every step",
                       "sharding": "batch (independent imagesets per rank, no data-path collective)",
                       "power": "this workload runs at the board power cap (see clocks: sw_power_cap, SM clock well below "
                                "max); the device-resident loop holds the cap continuously, while the copy phases of the e2e "
                                "loop let the clocks recover, so e2e can match or exceed value on the same box"},
            "model_tflops": flops_step * world * args.steps / ms_total / 1e9,
            "roofline": roofline,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": (b * l * s * s + b * l) * 4,
                    "d2h_bytes_per_step": b * 9 * s * s * 4, "ms_per_step": ms_e2e / args.steps,
                    "api": "HRNet.forward_host_submit / forward_host_wait (hrn_forward_host_submit): pinned host lrs/alphas "
                           "in, pinned host sr out, every step; two steps in flight",
                    "one_call_at_a_time": {"value": total_sets / ms_e2e_sync * 1e3, "ms_per_step": ms_e2e_sync / args.steps,
                                           "api": "HRNet.forward_host (hrn_forward_host), synchronous"}},
            "gpu_launches": launches,
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(l, s)
        else:
            line["cpu_baseline"] = None
        print(json.dumps(line), file=out, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
