"""Two launches of the pair GEMM at M = 4095, N = 5120, K = 13824 (256- then 224-wide tiles) for an `ncu --set full`
capture: profiles/r02_gemm_tile_ncu.csv."""
import sys, torch
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from goal_force_b200 import capi
M, N, K = 4095, 5120, 13824
a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
w = (torch.randn(N, K, device="cuda") * K ** -0.5).to(torch.bfloat16)
b = torch.zeros(N, device="cuda", dtype=torch.bfloat16)
out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
for bn in (256, 224):
    capi.gemm_tile_tuning(bn)
    capi.gemm(a, w, b, out=out)
torch.cuda.synchronize()
