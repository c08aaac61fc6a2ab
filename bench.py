#!/usr/bin/env python
"""bench.py -- denoise steps/s of the Goal Force A14B DiT forward on B200 (BASELINE.json metric).

    python bench.py --gpus 1 --steps 3 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
           bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...        # the reference's own PyTorch CPU path (oracle port) on the host cores

A "step" is one pass of the hot path: one `model_fn_wan_video` call = one A14B DiT forward of the high-noise expert in
goal-force mode (40 trunk blocks + 10-block ControlNet with non-zero zero-convs) at 81x480x832 -> 32,760 tokens
(BASELINE.json configs[1]). Weights are random-init of that architecture, inputs are synthetic latents (no network).
At N > 1 the token sequence is sharded over the N GPUs (Ulysses sequence parallel, strong scaling).
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

METRIC = "A14B DiT denoise steps/sec (81x480x832, 32760 tokens, goal-force ControlNet)"
UNIT = "steps/s"
FRAMES_LAT, H_LAT, W_LAT = 21, 60, 104          # 81 x 480 x 832 video -> latent grid; tokens = 21*30*52 = 32760
CONTROLNET_LAYERS = 10


def attn_dram_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE self-attention launch (gf_attn80x2_kernel, the kernel the
    library picks at this shape: L = 32760, 40 heads), parsed from the newest committed `ncu --set full` summary
    profiles/rNN_attn80x2_ncu.csv (units are in the header; older rounds: rNN_attn80_ncu.csv).
    Returns (bytes, source) or (None, reason)."""
    import csv
    import re
    cands = sorted((ROOT / "profiles").glob("r*_attn80x2_ncu.csv"), reverse=True) + \
        sorted((ROOT / "profiles").glob("r*_attn80_ncu.csv"), reverse=True)
    for path in cands:
        try:
            rows = list(csv.reader(path.read_text().splitlines()))
            head, row = rows[0], rows[1]
            total = 0.0
            for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                col = next(i for i, h in enumerate(head) if h.startswith(key))
                unit = re.search(r"\[(\w+)\]", head[col]).group(1).lower()
                mult = {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[unit]
                total += float(row[col]) * mult
            return total, f"profiles/{path.name}"
        except Exception:  # noqa: BLE001 - malformed file: try the next one
            continue
    return None, "no profiles/r*_attn80_ncu.csv"


def gpu_reference_context():
    """Context, not a target and not the reference arm: the reference algorithm as eager PyTorch + cuDNN attention on
    the same kind of GPU, measured by tools/gpu_oracle_baseline.py and committed under profiles/."""
    cands = sorted((ROOT / "profiles").glob("r*_gpu_eager_oracle.json"), reverse=True)
    for path in cands:
        try:
            d = json.loads(path.read_text())
            return {"what": d.get("what"), "ms_per_step": round(1000.0 * d["extrapolated_seconds_per_forward_50_blocks"], 1),
                    "timed_blocks": d.get("timed_blocks"), "full_forward_measured": d.get("full_forward_measured", False),
                    "source": f"profiles/{path.name}", "note": "committed measurement, not taken in this run"}
        except Exception:  # noqa: BLE001
            continue
    return None


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--layers", type=int, default=40, help="trunk blocks (40 = A14B; fewer only for debugging)")
    ap.add_argument("--controlnet-layers", type=int, default=CONTROLNET_LAYERS)
    ap.add_argument("--transport", default="peer", choices=["peer", "nccl"],
                    help="Ulysses exchange at N > 1: fused peer-memory stores (default) or NCCL all-to-all")
    ap.add_argument("--controlnet-stream", default="auto", choices=["auto", "on", "off"],
                    help="ControlNet branch on a second stream next to the trunk (auto: on for N > 1)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--breakdown", action="store_true", help="also print a per-kernel table to stderr")
    return ap.parse_args()


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return {"tflops": d.get("bf16_tflops_sustained", 1400.0), "tflops_burst": d.get("bf16_tflops", 1590.0),
                "hbm_gbs": d.get("hbm_gbs", 6650.0), "source": "measured"}
    return {"tflops": 1400.0, "tflops_burst": 1590.0, "hbm_gbs": 6650.0, "source": "fallback"}


def total_flops(L, d, ffn, ctx, layers):
    """Work executed per step: self-attention linears + attention, cross-attention q/o linears + attention, FFN.  The
    cross-attention K/V projections of the 512 context tokens are step-invariant and cached, so they are not counted."""
    per_block = 8 * L * d * d + 4 * L * L * d + 4 * L * d * d + 4 * L * ctx * d + 4 * L * d * ffn
    return per_block * layers


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx = float(r[2])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except (ValueError, IndexError):
                continue
        sm.sort()
        # median over the upper half of samples = clocks under load (idle samples at the edges drop out)
        load = sm[len(sm) // 2:] if sm else []
        med = load[len(load) // 2] if load else None
        return {"sm_mhz": med, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ CPU reference arm
def cpu_reference_sample(threads: int):
    """Times the reference algorithm (oracle port, bit-identical to the reference's PyTorch code on CPU) on ONE full
    A14B DiTBlock at the full config-2 length -- 32,760 tokens, all 40 heads: LayerNorms, q/k/v/o + RMSNorm + complex128
    RoPE, full-length self-attention, cross-attention against 512 context tokens, FFN -- exactly as SURVEY 8(d)
    prescribes (no token or head sub-sampling).  One step = 50 such blocks (40 trunk + 10 ControlNet; the patch embed,
    head and zero-conv GEMMs, < 1 % of the work, are left out), so steps/s = 1 / (50 x t_block), labelled extrapolated.
    bf16, torch CPU kernels, all host threads.  Returns (steps_per_s, seconds_measured, description)."""
    import torch
    from oracle import wan_dit_oracle as O
    torch.set_num_threads(threads)
    cfg = O.DiTConfig(**{**O.WAN22_I2V_A14B.__dict__, "num_layers": 1})
    L = FRAMES_LAT * (H_LAT // 2) * (W_LAT // 2)
    g = torch.Generator("cpu").manual_seed(0)
    sd = {}
    rn = lambda *s, scale=1.0: (torch.randn(*s, generator=g) * scale)  # noqa: E731
    O._random_block(sd, "blocks.0", cfg, rn,
                    lambda n, o, i: sd.update({n + ".weight": rn(o, i, scale=i ** -0.5), n + ".bias": rn(o, scale=0.02)}))
    sd = {k: v.to(torch.bfloat16) for k, v in sd.items()}
    x = torch.randn(1, L, cfg.dim, generator=g).bfloat16()
    ctx = torch.randn(1, 512, cfg.dim, generator=g).bfloat16()
    t_mod = torch.randn(1, 6, cfg.dim, generator=g).bfloat16()
    freqs = O.rope_freqs(cfg.head_dim, FRAMES_LAT, H_LAT // 2, W_LAT // 2, "cpu")
    with torch.no_grad():
        O.dit_block(sd, "blocks.0", x[:, :64], ctx, t_mod, freqs[:64], cfg)           # warm-up (thread pool, caches)
        t0 = time.perf_counter()
        O.dit_block(sd, "blocks.0", x, ctx, t_mod, freqs, cfg)
        t_block = time.perf_counter() - t0
    blocks = 40 + CONTROLNET_LAYERS
    t_step = blocks * t_block
    desc = (f"oracle port of the reference PyTorch CPU path, bf16, {threads} threads: ONE full A14B DiTBlock at "
            f"{L} tokens x 40 heads measured ({t_block:.2f}s), EXTRAPOLATED x{blocks} blocks to one step "
            f"(SURVEY 8(d)); embeddings/head/zero-convs (<1%) not included")
    return 1.0 / t_step, t_block, desc


def run_reference(args, emit):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    # each timed step is one bounded sample (one full-length DiTBlock); the warm-up is the small call inside the sample
    # (thread pool / allocator), W full-length warm-up blocks would only burn minutes of host time
    vals, secs, desc = [], 0.0, ""
    for _ in range(max(1, args.steps)):
        v, s, desc = cpu_reference_sample(threads)
        vals.append(v)
        secs += s
    value = len(vals) / sum(1.0 / v for v in vals)          # steps/s over the measured samples (harmonic mean)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1000.0 / value, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "impl": "reference",
            "config": workload_config(args, 1), "extrapolated": True, "seconds_measured": round(secs, 2),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": desc},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(json.dumps(line))


def workload_config(args, n):
    return {"workload": "configs[1]: Wan2.2 I2V A14B high-noise expert, one denoise step (= one DiT forward), Goal "
                        "Force mode (10-block ControlNet, target-force control latents), 81x480x832 = 32760 tokens",
            "tokens": FRAMES_LAT * (H_LAT // 2) * (W_LAT // 2), "trunk_blocks": args.layers,
            "controlnet_blocks": args.controlnet_layers, "controlnet_stream": args.controlnet_stream, "parallelism": f"ulysses_sp{n}_{args.transport}" if n > 1 else "single_gpu",
            "l2": "per-step working set (35 GB of weights + 3 GB of activations) is far larger than the 126 MB L2; "
                  "no explicit flush"}


# ------------------------------------------------------------------------------------------------ our arm
def run_ours(args, emit):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run for N > 1")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from goal_force_b200 import capi
    from goal_force_b200.pipeline import ParallelContext, ParallelLayout
    from goal_force_b200.synthetic import LazyRandomStateDict, synthetic_inputs
    from goal_force_b200.wan_dit import (ControlNetB200, DiTConfig, WAN22_I2V_A14B, WanModelB200, model_fn_wan_video)
    capi.load()
    cfg = DiTConfig(**{**WAN22_I2V_A14B.__dict__, "num_layers": args.layers})
    par = (ParallelContext(ParallelLayout(world_size=world, rank=rank, cfg_size=1), transport=args.transport)
           if world > 1 else None)
    sp = par.sp if par is not None else None
    dit = WanModelB200(cfg, LazyRandomStateDict(cfg, seed=0, device=dev), device=dev)
    cn = None
    if args.controlnet_layers > 0:
        cn = ControlNetB200(cfg, LazyRandomStateDict(cfg, seed=1, device=dev, controlnet_layers=args.controlnet_layers),
                            args.controlnet_layers, device=dev)
    host = synthetic_inputs(cfg, FRAMES_LAT, H_LAT, W_LAT, seed=1, device="cpu", pin=True, timestep=937.0)
    devin = {k: v.to(dev) for k, v in host.items()}
    L = FRAMES_LAT * (H_LAT // 2) * (W_LAT // 2)

    def step(inp):
        return model_fn_wan_video(dit=dit, controlnet=cn, latents=inp["latents"], timestep=inp["timestep"],
                                  context=inp["context"], y=inp["y"],
                                  control_signal_video_latents=inp["control_signal_video_latents"],
                                  sequence_parallel=sp,
                                  controlnet_stream={"auto": None, "on": True, "off": False}[args.controlnet_stream])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        out = step(devin)
    barrier()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    # ---- timed region 1: inputs resident in HBM
    capi.STATS.reset(timing=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        out = step(devin)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = capi.STATS.launches
    kern = capi.STATS.summary()
    capi.STATS.reset(timing=False)
    # ---- timed region 2: end to end through the public call with host buffers (H2D of every input, D2H of result)
    e2e = None
    h2d = sum(v.numel() * v.element_size() for v in host.values())
    d2h = out.numel() * out.element_size()
    if not args.no_e2e:
        res_host = torch.empty(out.shape, dtype=out.dtype).pin_memory()
        bufs = {k: torch.empty_like(v, device=dev) for k, v in host.items()}

        def e2e_step():
            for k, v in host.items():
                bufs[k].copy_(v, non_blocking=True)
            o = step(bufs)
            res_host.copy_(o, non_blocking=True)
            torch.cuda.current_stream().synchronize()

        # Every copy into `bufs` bumps the tensors' versions, so each e2e step rebuilds the step-invariant caches
        # (text embedding, 50 x cross-attention K|V, ControlNet tokens): that work stays inside the timed region.  What
        # the untimed pass below removes is the ONE-TIME cost of the first such step: ~50 fresh cudaMalloc calls for
        # the second generation of cache entries (each one drains the launch queue), 20-360 ms spread over K steps.
        for _ in range(min(args.warmup, 1)):
            e2e_step()
        barrier()
        e0.record()
        for _ in range(args.steps):
            e2e_step()
        e1.record()
        barrier()
        e2e = e0.elapsed_time(e1)
    clk = clocks.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms, e2e or 0.0], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e = float(t[0]), (float(t[1]) if e2e is not None else None)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    pk = peaks()
    ms_step = ms / args.steps
    value = 1000.0 / ms_step
    blocks = args.layers + (args.controlnet_layers if cn is not None else 0)
    flops = total_flops(L, cfg.dim, cfg.ffn_dim, 512, blocks) + (2.0 * L * cfg.dim * cfg.dim * args.controlnet_layers
                                                                 if cn is not None else 0)
    att = kern.get("attention_self", {"launches": 0, "ms": 0.0, "work": 0.0})
    roof = None
    traffic, traffic_src = attn_dram_traffic()
    if att["launches"]:
        ach = att["work"] / att["ms"] / 1e9
        roof = {"kernel": "gf_attn80x2_kernel (self-attention, tcgen05 flash attention, CTA pair)" if world == 1 else
                          "self-attention (gf_attn80x2_kernel / gf_attn80_kernel by shape)", "bound": "tensor",
                "achieved": round(ach, 1), "peak": pk["tflops"], "unit": "TFLOP/s", "frac": round(ach / pk["tflops"], 4),
                "peak_source": f"{pk['source']} (cuBLAS bf16 sustained, MEASURED_PEAKS.json)",
                # dram__bytes_read.sum + dram__bytes_write.sum of one launch at this shape, parsed from the committed
                # ncu --set full summary (algorithmic Q+K+V+O = 1.342e9 bytes)
                "traffic": traffic if (world == 1 and args.layers > 0) else None,
                "traffic_unit": f"bytes/launch (ncu, {traffic_src})",
                "algorithmic_bytes_per_launch": 4.0 * L * cfg.dim * 2,
                "launches": att["launches"], "avg_ms": round(att["ms"] / att["launches"], 4),
                "share_of_step": round(att["ms"] / ms, 4),
                "flops_per_launch": att["work"] / att["launches"]}
    kernels = {}
    for tag, d in sorted(kern.items(), key=lambda kv: -kv[1]["ms"]):
        ent = {"launches": d["launches"], "ms_per_step": round(d["ms"] / args.steps, 3),
               "share": round(d["ms"] / ms, 4)}
        if tag in ("gemm", "attention_self", "attention_cross"):
            ent["tflops"] = round(d["work"] / d["ms"] / 1e9, 1)
            ent["frac_of_peak"] = round(ent["tflops"] / pk["tflops"], 4)
        else:
            ent["gbs"] = round(d["work"] / d["ms"] / 1e6, 1)
            ent["frac_of_hbm"] = round(ent["gbs"] / pk["hbm_gbs"], 4)
        kernels[tag] = ent
    line = {"metric": METRIC, "value": round(value, 5), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms_step, 3), "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "impl": "ours",
            "config": workload_config(args, world),
            "model_tflops_per_step": round(flops / 1e12, 1),
            "achieved_tflops_per_gpu": round(flops / 1e12 / (ms_step / 1e3) / world, 1),
            "frac_of_dense_bf16_spec_2250": round(flops / 1e12 / (ms_step / 1e3) / world / 2250.0, 4),
            "frac_of_measured_sustained": round(flops / 1e12 / (ms_step / 1e3) / world / pk["tflops"], 4),
            "roofline": roof, "kernels": kernels, "gpu_launches": launches, "clocks": clk}
    if e2e is not None:
        line["e2e"] = {"value": round(1000.0 / (e2e / args.steps), 5), "unit": UNIT, "h2d_bytes_per_step": h2d,
                       "d2h_bytes_per_step": d2h, "ms_per_step": round(e2e / args.steps, 3)}
    ctx_ref = gpu_reference_context()
    if ctx_ref is not None and world == 1:
        line["gpu_reference_context"] = ctx_ref
    if world == 1 and not args.no_cpu_baseline:
        v, _, desc = cpu_reference_sample(os.cpu_count() or 1)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port", "sample": desc}
    if args.breakdown:
        for tag, ent in kernels.items():
            print(f"  {tag:16s} {ent}", file=sys.stderr)
    emit(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


class _StdoutToStderr:
    """Rank 0's stdout must carry ONE JSON line, but NCCL prints its version banner (and libraries may print other
    things) straight to file descriptor 1.  While active, fd 1 points at stderr; `emit` writes to the real stdout."""

    def __enter__(self):
        sys.stdout.flush()
        self._real = os.dup(1)
        os.dup2(2, 1)
        return self

    def emit(self, text: str) -> None:
        os.write(self._real, (text + "\n").encode())

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self._real, 1)
        os.close(self._real)
        return False


def main():
    args = parse()
    with _StdoutToStderr() as out:
        if args.impl == "reference":
            run_reference(args, out.emit)
        else:
            run_ours(args, out.emit)


if __name__ == "__main__":
    main()
