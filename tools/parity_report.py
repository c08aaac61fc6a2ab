#!/usr/bin/env python
"""Parity numbers on the GPU in one table (the assertions live in tests/; this prints the measured values).

    python tools/parity_report.py > gpurun_out/parity_report.md

For every case: relL2(ours, oracle fp32), relL2(oracle bf16, oracle fp32) (= the reference's own bf16 noise on this
device) and relL2(ours, oracle bf16); the oracle is oracle/wan_dit_oracle.py run on the same GPU.
"""
from __future__ import annotations

import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

from oracle import wan_dit_oracle as O  # noqa: E402


def forward_case(name, cfg, shape, n_cn, seeds=(0, 1, 2), timestep=900.0):
    from goal_force_b200.wan_dit import ControlNetB200, DiTConfig, WanModelB200, model_fn_wan_video
    sd = O.random_state_dict(cfg, seed=seeds[0])
    csd = O.random_controlnet_state_dict(cfg, n_cn, seed=seeds[1]) if n_cn else None
    inp = O.synthetic_inputs(cfg, *shape, seed=seeds[2], timestep=timestep)
    outs = []
    for dt in (torch.float32, torch.bfloat16):
        s = {k: v.to("cuda", dt) for k, v in sd.items()}
        i = {k: v.to("cuda", dt) for k, v in inp.items()}
        kw = {}
        if csd is not None:
            kw = dict(controlnet_sd={k: v.to("cuda", dt) for k, v in csd.items()},
                      control_signal_video_latents=i["control_signal_video_latents"], controlnet_num_layers=n_cn)
        with torch.no_grad():
            outs.append(O.model_fn(s, cfg, i["latents"], i["timestep"], i["context"], y=i.get("y"), **kw))
        del s
    pc = DiTConfig(**cfg.__dict__)
    bf = {k: v.to("cuda", torch.bfloat16) for k, v in inp.items()}
    kw = {}
    if csd is not None:
        kw = dict(controlnet=ControlNetB200(pc, csd, n_cn), control_signal_video_latents=bf["control_signal_video_latents"])
    out = model_fn_wan_video(dit=WanModelB200(pc, sd), latents=bf["latents"], timestep=bf["timestep"],
                             context=bf["context"], y=bf.get("y"), **kw)
    L = shape[0] * (shape[1] // 2) * (shape[2] // 2)
    print(f"| {name} | {L} | {O.rel_l2(out, outs[0]):.3e} | {O.rel_l2(outs[1], outs[0]):.3e} | "
          f"{O.rel_l2(out, outs[1]):.3e} | {O.cosine(out, outs[0]):.6f} |", flush=True)


def main():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    from goal_force_b200 import capi
    capi.load()
    print(f"# Parity report ({torch.cuda.get_device_name(0)}, torch {torch.__version__})\n")
    print("Tolerance used by the tests: relL2(ours, fp32) <= max(1e-2, 1.1 x relL2(oracle bf16, fp32)).\n")
    print("| case | tokens | ours vs oracle fp32 | oracle bf16 vs fp32 | ours vs oracle bf16 | cosine vs fp32 |")
    print("|---|---:|---:|---:|---:|---:|")
    forward_case("configs[0]: Wan2.1-T2V-1.3B shape, 30 blocks, 17x240x416", O.WAN21_T2V_1_3B, (5, 30, 52), 0)
    a14 = O.DiTConfig(**{**O.WAN22_I2V_A14B.__dict__, "num_layers": 2})
    forward_case("A14B widths, 2 blocks + 1 ControlNet block, latent 4x40x52", a14, (4, 40, 52), 1, seeds=(2, 3, 4), timestep=990.0)
    a14_4 = O.DiTConfig(**{**O.WAN22_I2V_A14B.__dict__, "num_layers": 4})
    forward_case("A14B widths, 4 blocks + 2 ControlNet blocks, latent 8x40x52", a14_4,
                 (8, 40, 52), 2, seeds=(5, 6, 7), timestep=937.0)
    forward_case("A14B widths, 2 blocks + 1 ControlNet block, latent 21x60x104 (81x480x832, config 2 length)", a14, (21, 60, 104), 1,
                 seeds=(8, 9, 10), timestep=937.0)


if __name__ == "__main__":
    main()
