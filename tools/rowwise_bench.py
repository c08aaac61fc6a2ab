"""Streaming kernels of the VAE at their full-size shapes: achieved HBM bandwidth (algorithmic bytes / CUDA-event time)."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from goal_force_b200 import capi  # noqa: E402


def timed(fn, iters=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    for (T, H, W, C) in ((81, 480, 832, 96), (81, 240, 416, 192), (41, 120, 208, 384)):
        x = torch.randn(T, H, W, C, device="cuda").to(torch.bfloat16)
        g = torch.ones(C, device="cuda", dtype=torch.bfloat16)
        y = torch.empty_like(x)
        ms = timed(lambda: capi.vae_rmsnorm(x, g, silu=True, out=y))
        print(f"rmsnorm {T}x{H}x{W}x{C}: {ms:.3f} ms  {2 * x.numel() * 2 / ms / 1e6:.0f} GB/s", flush=True)
        del x, y
    for (T, H, W, C, temporal) in ((81, 240, 416, 192, False), (41, 120, 208, 384, True)):
        x = torch.randn(T, H, W, C, device="cuda").to(torch.bfloat16)
        if temporal:
            rest = torch.randn(T - 1, H, W, 2 * C, device="cuda").to(torch.bfloat16)
            F = 2 * T - 1
        else:
            rest, F = None, T
        out = torch.empty(F, 2 * H, 2 * W, C, device="cuda", dtype=torch.bfloat16)
        ms = timed(lambda: capi.vae_upsample2x(x, rest, F, H, W, C, out=out))
        print(f"upsample2x {F}x{H}x{W}x{C} temporal={temporal}: {ms:.3f} ms  {5 * F * H * W * C * 2 / ms / 1e6:.0f} GB/s", flush=True)
        del x, out, rest


if __name__ == "__main__":
    main()
