"""Per-call device time of shift_cPSNR_argmax right after start-up (no clock ramp): shows allocator / clock transients."""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pynvml
import highres_net_b200 as hb
pynvml.nvmlInit(); nv = pynvml.nvmlDeviceGetHandleByIndex(0)
dev = torch.device("cuda:0")
for n in (512, 32):
    sr, hr, hm = torch.rand(n, 384, 384, device=dev), torch.rand(n, 384, 384, device=dev), (torch.rand(n, 384, 384, device=dev) > 0.1).float()
    for generic in (0, 1):
        hb.scoring_debug_set("cpsnr_generic", generic)
        torch.cuda.synchronize(); time.sleep(0.5)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(31)]
        t0 = time.time()
        ev[0].record()
        for i in range(30):
            hb.shift_cPSNR_argmax(sr, hr, hm)
            ev[i + 1].record()
        t_enq = time.time() - t0
        torch.cuda.synchronize()
        ms = [round(ev[i].elapsed_time(ev[i + 1]), 3) for i in range(30)]
        print(json.dumps({"n": n, "generic": generic, "enqueue_ms_per_call": round(t_enq / 30 * 1e3, 3), "sm_mhz_after": pynvml.nvmlDeviceGetClockInfo(nv, pynvml.NVML_CLOCK_SM), "ms": ms}), flush=True)
hb.scoring_debug_set("cpsnr_generic", 0)
