"""Differential fuzz of the two N = 7 Lanczos kernels (TMA pipelines vs register window) on random shapes and paddings: they
must agree bit for bit.   python tools/lanczos_fuzz.py [cases]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import highres_net_b200 as hb
dev = torch.device("cuda:0")
rng = np.random.RandomState(99)
n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 100
bad = 0
for case in range(n_cases):
    h = int(rng.randint(1, 140)) if rng.rand() < 0.7 else int(rng.randint(140, 420))
    w = 4 * int(rng.randint(2, 36)) if rng.rand() < 0.7 else 4 * int(rng.randint(36, 110))
    nb, c = int(rng.randint(1, 3)), int(rng.randint(1, 7))
    p = int(rng.randint(0, min(h, w, 8)))
    img = torch.from_numpy(rng.rand(nb, c, h, w).astype(np.float32)).to(dev)
    sh = torch.from_numpy(rng.uniform(-1.5, 1.5, size=(c, 2)).astype(np.float32)).to(dev)
    hb.scoring_debug_set("lanczos_scalar", 0); a = hb.lanczos_shift(img, sh, p=p)
    hb.scoring_debug_set("lanczos_scalar", 1); b = hb.lanczos_shift(img, sh, p=p)
    hb.scoring_debug_set("lanczos_scalar", 0)
    ok = bool(torch.equal(a, b))
    bad += not ok
    if not ok or case % 10 == 0:
        print(f"case {case:3d} nb={nb} c={c} h={h:3d} w={w:3d} p={p} identical={ok} max diff {float((a - b).abs().max()):.2e}", flush=True)
print(f"failures {bad} of {n_cases}")
sys.exit(1 if bad else 0)
