"""Per-shape timing of gf_conv3d_cl_bf16 at the Wan VAE's full-size layer shapes (81 x 480 x 832 decode), device-timed.
Usage: python tools/conv_bench.py [--shapes s3,s2,...] [--iters 3] [--fused]"""
import argparse
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

SHAPES = {
    # name: (Cin, Cout, (T, H, W), kernel, pad)
    "s3": (96, 96, (81, 480, 832), (3, 3, 3), (2, 1, 1)),
    "s2": (192, 192, (81, 240, 416), (3, 3, 3), (2, 1, 1)),
    "s1": (384, 384, (41, 120, 208), (3, 3, 3), (2, 1, 1)),
    "s0": (384, 384, (21, 60, 104), (3, 3, 3), (2, 1, 1)),
    "up2": (192, 96, (81, 480, 832), (1, 3, 3), (0, 1, 1)),
    "up1": (384, 192, (81, 240, 416), (1, 3, 3), (0, 1, 1)),
    "tile_s3": (96, 96, (81, 240, 416), (3, 3, 3), (2, 1, 1)),
    "head": (96, 3, (81, 480, 832), (3, 3, 3), (2, 1, 1)),
    "enc1": (8, 96, (81, 480, 832), (3, 3, 3), (2, 1, 1)),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shapes", default="s3,s2,s1,s0,up2,up1,tile_s3,head,enc1")
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--fused", action="store_true", help="residual + fused next-layer norm epilogue (Cout <= 256)")
    ap.add_argument("--epi", default="", help="plain | norm (Y2 only) | rawnorm (Y + Y2) | res (R + Y) | resnorm (R + Y + Y2); overrides --fused")
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    from goal_force_b200 import capi
    res = {}
    for name in a.shapes.split(","):
        cin, cout, (T, H, W), kernel, pad = SHAPES[name]
        taps = kernel[0] * kernel[1] * kernel[2]
        x = torch.randn(T, H, W, cin, device="cuda").to(torch.bfloat16)
        w = (torch.randn(cout, taps * cin, device="cuda") * (taps * cin) ** -0.5).to(torch.bfloat16)
        cs = (cout + 7) // 8 * 8
        b = torch.zeros(cs, dtype=torch.bfloat16, device="cuda")
        ncthw = cout < 8
        epi = a.epi or ("resnorm" if a.fused else "plain")
        if cout > 256 or ncthw:
            epi = "plain"
        fused = epi != "plain"
        want_norm, want_res, want_raw = epi in ("norm", "resnorm", "rawnorm"), epi in ("res", "resnorm"), epi != "norm"
        gamma = torch.ones(cs, dtype=torch.bfloat16, device="cuda") if want_norm else None
        resid = torch.randn(T, H, W, cs, device="cuda").to(torch.bfloat16) if want_res else None
        y = torch.empty((cout, T, H, W) if ncthw else (T, H, W, cs), dtype=torch.bfloat16, device="cuda") if want_raw else None
        yn = torch.empty((T, H, W, cs), dtype=torch.bfloat16, device="cuda") if want_norm else None

        def run():
            capi.conv3d_cl(x, w, b, kernel=kernel, pad=pad, out=y, residual=resid, norm_out=yn, gamma=gamma, cout=cout,
                           ncthw=ncthw, want_raw=want_raw)
        run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.iters):
            run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / max(a.iters, 1)
        flops = 2.0 * T * H * W * cout * taps * cin
        res[name] = {"ms": round(ms, 3), "tflops": round(flops / ms / 1e9, 1), "gflop": round(flops / 1e9, 1),
                     "epi": epi}
        print(name, res[name], flush=True)
        del x, y, yn, resid
        torch.cuda.empty_cache()
    if a.out:
        Path(a.out).write_text(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
