#!/usr/bin/env python
"""Measurements for the BASELINE.json configs that bench.py does not cover (bench.py is configs[1], one forward).

    torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/bench_configs.py --config sample --steps 40
    torchrun ... tools/bench_configs.py --config sample --cfg-parallel          # configs[3] layout: cfg 2 x ulysses N/2
    torchrun ... tools/bench_configs.py --config long --frames 81               # configs[4]: 81x720x1280 = 75,600 tokens
    torchrun ... tools/bench_configs.py --config long --frames 121              # 121x720x1280 = 111,600 tokens

  sample : configs[2]/[3] -- full two-expert A14B sampling: `--steps` denoise steps with CFG 5.0, shift 5.0, expert
           switch at t < 875, goal-force ControlNet (10 blocks) on the high-noise expert, never-loaded (all-zero,
           skipped) ControlNet on the low-noise expert as in the shipped inference script; 81x480x832.
  long   : configs[4] -- attention-bound stress, one forward of the high-noise expert + ControlNet at 720x1280.
  direct : configs[3] -- Direct Force mode (projectile force + mass channels), a batch of 4 CSV rows through
           goal_force_b200.jobs.BatchDriver: control videos synthesised bit-exactly on the host (overlapped with the
           previous row's denoising), encoded by a synthetic stand-in for the VAE, then the full two-expert sampling.
           Layout: --replicas R --cfg-parallel, e.g. 8 GPUs as cfg 2 x ulysses 4 (R = 1, rows one after the other)
           or as 4 replicas x cfg 2 (R = 4, one row per replica).  Reports videos/s and denoise steps/s.
Random-init weights, synthetic latents. Prints one JSON line on rank 0. Timing: CUDA events, max over ranks.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", required=True, choices=["sample", "long", "direct"])
    ap.add_argument("--replicas", type=int, default=1)
    ap.add_argument("--rows", type=int, default=4)
    ap.add_argument("--steps", type=int, default=40, help="sample: denoise steps; long: timed forwards")
    ap.add_argument("--frames", type=int, default=81)
    ap.add_argument("--cfg-parallel", action="store_true")
    ap.add_argument("--transport", default="peer", choices=["peer", "nccl"])
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    from goal_force_b200 import capi
    from goal_force_b200.pipeline import GoalForceDenoiser, ParallelContext, ParallelLayout, generate_noise
    from goal_force_b200.synthetic import LazyRandomStateDict, synthetic_inputs
    from goal_force_b200.wan_dit import ControlNetB200, WAN22_I2V_A14B as cfg, WanModelB200, model_fn_wan_video
    capi.load()
    cfg_size = 2 if (args.cfg_parallel and world > 1) else 1
    par = (ParallelContext(ParallelLayout(world_size=world, rank=rank, cfg_size=cfg_size, replicas=args.replicas),
                           transport=args.transport) if world > 1 else None)

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def lazy(seed, **kw):
        return LazyRandomStateDict(cfg, seed=seed, device=dev, **kw)

    dit = WanModelB200(cfg, lazy(0), device=dev)
    cn = ControlNetB200(cfg, lazy(1, controlnet_layers=10), 10, device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def elapsed_max_ms():
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t[0])
        return ms

    if args.config == "long":
        fl, hl, wl = (args.frames - 1) // 4 + 1, 720 // 8, 1280 // 8
        L = fl * (hl // 2) * (wl // 2)
        inp = synthetic_inputs(cfg, fl, hl, wl, seed=1, device=dev, timestep=937.0)
        sp = par.sp if par is not None else None

        def fwd():
            return model_fn_wan_video(dit=dit, controlnet=cn, latents=inp["latents"], timestep=inp["timestep"],
                                      context=inp["context"], y=inp["y"],
                                      control_signal_video_latents=inp["control_signal_video_latents"],
                                      sequence_parallel=sp)
        fwd()
        sync()
        e0.record()
        for _ in range(args.steps):
            fwd()
        e1.record()
        sync()
        ms = elapsed_max_ms() / args.steps
        d, ffn = cfg.dim, cfg.ffn_dim
        flops = 50 * (8 * L * d * d + 4 * L * L * d + 4 * L * d * d + 4 * 512 * d * d + 4 * L * 512 * d + 4 * L * d * ffn) \
            + 10 * 2.0 * L * d * d
        line = {"config": f"configs[4]: A14B DiT forward + 10-block ControlNet at {args.frames}x720x1280", "tokens": L,
                "ms_per_forward": ms, "forwards_per_s": 1000.0 / ms, "tflops_per_gpu": flops / 1e12 / (ms / 1e3) / world}
    elif args.config == "direct":
        from goal_force_b200 import jobs as J
        golden = json.loads((ROOT / "tests" / "golden" / "control_channels.json").read_text())
        names = ["_pendulum", "_toycar", "_cantaloupes", "_paw_tool2", "_golf", "_tennis", "_soccer_tool", "_pool_tool"]
        rows = [J.direct_force_row(golden["rows"][names[i % len(names)]], 250.0, 37.0, 2.5) for i in range(args.rows)]
        dit2 = WanModelB200(cfg, lazy(2), device=dev)
        cn2 = ControlNetB200(cfg, lazy(3, controlnet_layers=10, zero_convs=True), 10, device=dev)
        inp = synthetic_inputs(cfg, 21, 60, 104, seed=1, device=dev)
        ctx_n = torch.randn(1, 512, cfg.text_dim, generator=torch.Generator("cpu").manual_seed(7)).to(dev, torch.bfloat16)
        den = GoalForceDenoiser(dit, dit2, cn, cn2, parallel=par)
        cond = lambda row: dict(context_posi=inp["context"], context_nega=ctx_n, y=inp["y"])  # noqa: E731
        mk = lambda steps: J.BatchDriver(den, J.synthetic_control_encoder(dev), cond, parallel=par,  # noqa: E731
                                         num_inference_steps=steps, device=dev)
        mk(2).run(rows[:max(1, args.replicas)])                     # warm-up: both experts, caches, exchange buffers
        sync()
        e0.record()
        out = mk(args.steps).run(rows, seed=0)
        e1.record()
        sync()
        ms = elapsed_max_ms()
        assert all(not torch.isnan(j.result).any() for j in out)
        import hashlib

        def digest(t):       # the SHA-256 prefix recorded in tests/golden/control_channels.json
            return hashlib.sha256(t.contiguous().view(torch.int16).numpy().tobytes()).hexdigest()[:16]
        ok = all(digest(j.control_video) == golden["direct_force"][names[j.index % len(names)]]
                 for j in out if names[j.index % len(names)] in golden["direct_force"])
        line = {"config": f"configs[3]: Direct Force, {args.rows} CSV rows, {args.steps} steps each, CFG 5.0, 81x480x832",
                "tokens": 32760, "seconds_total": ms / 1e3, "videos_per_s": args.rows / (ms / 1e3),
                "denoise_steps_per_s": args.rows * args.steps / (ms / 1e3), "rows_on_rank0": [j.index for j in out],
                "control_videos_bit_exact_vs_reference_digests": ok}
    else:
        dit2 = WanModelB200(cfg, lazy(2), device=dev)
        cn2 = ControlNetB200(cfg, lazy(3, controlnet_layers=10, zero_convs=True), 10, device=dev)   # F6: exact no-op
        assert cn2.is_noop
        inp = synthetic_inputs(cfg, 21, 60, 104, seed=1, device=dev)
        ctx_n = torch.randn(1, 512, cfg.text_dim, generator=torch.Generator("cpu").manual_seed(7)).to(dev, torch.bfloat16)
        noise = generate_noise(tuple(inp["latents"].shape), seed=0, device=dev)
        den = GoalForceDenoiser(dit, dit2, cn, cn2, parallel=par)
        used = []
        den(noise, inp["context"], ctx_n, y=inp["y"], control_latents=inp["control_signal_video_latents"],
            num_inference_steps=2, cfg_scale=5.0, sigma_shift=5.0, switch_DiT_boundary=0.9)     # warm-up: t = 1000 -> high-noise, 833 -> low-noise expert
        sync()
        e0.record()
        out = den(noise, inp["context"], ctx_n, y=inp["y"], control_latents=inp["control_signal_video_latents"],
                  num_inference_steps=args.steps, cfg_scale=5.0, sigma_shift=5.0,
                  callback=lambda i, t, lat: used.append(0 if float(t) >= 875 else 1))
        e1.record()
        sync()
        ms = elapsed_max_ms()
        assert not torch.isnan(out).any()
        line = {"config": f"configs[2]: two-expert A14B sampling, {args.steps} steps, CFG 5.0, 81x480x832", "tokens": 32760,
                "seconds_per_video_denoise": ms / 1e3, "denoise_steps_per_s": args.steps / (ms / 1e3),
                "high_noise_steps": used.count(0), "low_noise_steps": used.count(1)}
    line.update(n_gpus=world, layout=f"{args.replicas} replica(s) x cfg{cfg_size} x ulysses{world // cfg_size // args.replicas}"
                + (f" ({args.transport})" if world > 1 else ""),
                dtype="bf16", data="synthetic, random-init weights")
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
