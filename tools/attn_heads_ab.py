import sys, torch
sys.path.insert(0, '.')
from goal_force_b200 import capi
from tools.gpu_check import timed
L, d = 32760, 5120
qkv = torch.randn(L, 3 * d, device="cuda").bfloat16()
for h in (5, 10, 20):
    o = torch.empty(L, h * 128, device="cuda", dtype=torch.bfloat16)
    fl = 4.0 * L * L * h * 128
    res = {}
    for rnd in range(2):
        for impl in (80, 160):
            capi.attention_tuning(impl, 0)
            run = lambda: capi.attention(qkv[:, :h * 128], qkv[:, d:d + h * 128], qkv[:, 2 * d:2 * d + h * 128], h, out=o)
            ms = timed(run, iters=30, warmup=3)
            res.setdefault(impl, []).append(round(fl / ms / 1e9, 1))
    print("heads", h, res, flush=True)
