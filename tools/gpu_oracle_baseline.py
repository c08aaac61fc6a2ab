#!/usr/bin/env python
"""Context number (reported, not a target): the reference algorithm as eager PyTorch on the SAME B200 -- the oracle
port (identical torch ops to diffsynth's WanModel / goal-force model_fn; attention through
F.scaled_dot_product_attention, i.e. whatever fused kernel torch picks on sm_100) -- at config 2:
A14B widths, 32,760 tokens, bf16.  Default: the FULL forward bench.py times (40 trunk + 10 ControlNet blocks, 35 GB
of bf16 parameters resident, the same LazyRandomStateDict values our arm uses), sustained over --iters forwards, so
the number includes power throttling exactly like ours.  --layers / --controlnet-layers time a slice instead.

    python tools/gpu_oracle_baseline.py [--layers 40] [--controlnet-layers 10] [--iters 3] > profiles/rNN_gpu_eager_oracle.json
"""
from __future__ import annotations

import argparse
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--layers", type=int, default=40)
    ap.add_argument("--controlnet-layers", type=int, default=10)
    ap.add_argument("--iters", type=int, default=3)
    args = ap.parse_args()
    import torch
    from oracle import wan_dit_oracle as O
    from goal_force_b200.synthetic import LazyRandomStateDict
    from goal_force_b200.wan_dit import DiTConfig
    dev = torch.device("cuda", 0)
    cfg = O.DiTConfig(**{**O.WAN22_I2V_A14B.__dict__, "num_layers": args.layers})
    pcfg = DiTConfig(**cfg.__dict__)
    lazy = LazyRandomStateDict(pcfg, seed=0, device=dev)
    names = ["patch_embedding.weight", "patch_embedding.bias", "head.head.weight", "head.head.bias", "head.modulation",
             "time_projection.1.weight", "time_projection.1.bias"]
    names += [f"{m}.{i}.{p}" for m in ("text_embedding", "time_embedding") for i in (0, 2) for p in ("weight", "bias")]
    blk = [f"{a}.{q}.{p}" for a in ("self_attn", "cross_attn") for q in "qkvo" for p in ("weight", "bias")]
    blk += [f"{a}.norm_{q}.weight" for a in ("self_attn", "cross_attn") for q in "qk"]
    blk += ["norm3.weight", "norm3.bias", "ffn.0.weight", "ffn.0.bias", "ffn.2.weight", "ffn.2.bias", "modulation"]
    sd = {n: lazy[n] for n in names}
    for i in range(args.layers):
        sd.update({f"blocks.{i}.{n}": lazy[f"blocks.{i}.{n}"] for n in blk})
    ncn = args.controlnet_layers
    clazy = LazyRandomStateDict(pcfg, seed=1, device=dev, controlnet_layers=ncn)
    csd = {n: clazy[n] for n in ("controlnet_patch_embedding.patch_embedding.weight",
                                 "controlnet_patch_embedding.patch_embedding.bias")}
    for i in range(ncn):
        csd.update({f"controlnet_zero_convs_after.{i}.{p}": clazy[f"controlnet_zero_convs_after.{i}.{p}"]
                    for p in ("weight", "bias")})
        csd.update({f"controlnet_dit.blocks.{i}.{n}": clazy[f"controlnet_dit.blocks.{i}.{n}"] for n in blk})
    inp = {k: v.to(dev, torch.bfloat16) for k, v in O.synthetic_inputs(cfg, 21, 60, 104, seed=1, timestep=937.0).items()}

    def fwd():
        with torch.no_grad():
            return O.model_fn(sd, cfg, inp["latents"], inp["timestep"], inp["context"], y=inp["y"], controlnet_sd=csd,
                              control_signal_video_latents=inp["control_signal_video_latents"], controlnet_num_layers=ncn)

    fwd()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.iters):
        fwd()
    e1.record()
    torch.cuda.synchronize()
    sec = e0.elapsed_time(e1) / 1e3 / args.iters
    blocks = args.layers + ncn
    per_block = sec / blocks        # embeddings / head are < 1 % of a block
    full = per_block * 50
    print(json.dumps({"what": "oracle port (eager PyTorch bf16, SDPA attention) on this GPU, config-2 shape",
                      "timed_blocks": blocks, "trunk_blocks": args.layers, "controlnet_blocks": ncn, "iters": args.iters,
                      "full_forward_measured": blocks == 50, "seconds": sec, "seconds_per_block": per_block,
                      "extrapolated_seconds_per_forward_50_blocks": full, "steps_per_s": 1.0 / full,
                      "torch": torch.__version__, "gpu": torch.cuda.get_device_name(0)}))


if __name__ == "__main__":
    main()
