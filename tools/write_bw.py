"""Write-only and read-only HBM bandwidth next to the copy figure of MEASURED_PEAKS.json (conv_init writes 1.07 GB and reads
almost nothing, so its floor is the write-only number)."""
import torch, json
dev = torch.device("cuda:0")
n = 1 << 29                                   # 1 GiB of bf16
a = torch.empty(n, dtype=torch.bfloat16, device=dev); b = torch.empty_like(a)
def t(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
ms_fill = t(lambda: a.zero_())
ms_copy = t(lambda: b.copy_(a))
ms_read = t(lambda: a.view(torch.int16).max())
print(json.dumps({"write_only_GBps": 2 * n / ms_fill / 1e6, "copy_GBps_rw": 4 * n / ms_copy / 1e6, "read_only_GBps(max reduce)": 2 * n / ms_read / 1e6}))
