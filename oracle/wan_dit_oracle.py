"""CPU/torch restatement of the Goal Force denoising forward -- TEST INFRASTRUCTURE ONLY.

This file is the parity oracle for goal_force_b200. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import it; the product package never does.

It restates, with plain torch functional ops on whatever device/dtype the caller passes, the algorithm of
  diffsynth/models/wan_video_dit.py           (WanModel, DiTBlock, SelfAttention, CrossAttention, RMSNorm, Head, RoPE)
  src/goal_force/wan_video_new.py:1349-1591   (model_fn_wan_video incl. the goal-force ControlNet branch)
  src/goal_force/wan_video_new.py:40-117      (ControlNet modules)
Weights are a flat dict with the reference's state_dict key names, so a reference module's state_dict() can be fed
in unchanged. Pinning: oracle/gen_golden.py runs the real reference (imported from /root/reference) and this file on
the same seeded weights/inputs; tests/test_oracle_cpu.py checks the committed vectors in tests/golden/. On identical
device/dtype the two are bit-identical because they issue the same torch ops in the same order (the attention goes
through F.scaled_dot_product_attention, the branch the reference takes when flash-attn is absent,
wan_video_dit.py:55-60).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch
import torch.nn.functional as F


@dataclass(frozen=True)
class DiTConfig:
    """kwargs tables of the reference: wan_video_dit.py:502-514 (1.3B) and :703-718 (Wan2.2 I2V A14B)."""
    dim: int
    in_dim: int
    ffn_dim: int
    out_dim: int
    text_dim: int
    freq_dim: int
    eps: float
    num_heads: int
    num_layers: int
    patch_size: tuple = (1, 2, 2)

    @property
    def head_dim(self) -> int:
        return self.dim // self.num_heads


WAN21_T2V_1_3B = DiTConfig(dim=1536, in_dim=16, ffn_dim=8960, out_dim=16, text_dim=4096, freq_dim=256, eps=1e-6,
                           num_heads=12, num_layers=30)
WAN22_I2V_A14B = DiTConfig(dim=5120, in_dim=36, ffn_dim=13824, out_dim=16, text_dim=4096, freq_dim=256, eps=1e-6,
                           num_heads=40, num_layers=40)


# ----------------------------------------------------------------------------------------------------------- pieces
def sinusoidal_embedding_1d(dim: int, position: torch.Tensor) -> torch.Tensor:
    """wan_video_dit.py:68-72 -- float64 angles, result cast back to position.dtype (bf16 in the pipeline)."""
    half = dim // 2
    omega = torch.pow(10000, -torch.arange(half, dtype=torch.float64, device=position.device).div(half))
    ang = torch.outer(position.type(torch.float64), omega)
    return torch.cat([torch.cos(ang), torch.sin(ang)], dim=1).to(position.dtype)


def rope_axis_table(dim: int, end: int = 1024, theta: float = 10000.0) -> torch.Tensor:
    """wan_video_dit.py:83-89 -- complex128 e^{i pos theta^(-2j/dim)}, shape (end, dim/2)."""
    inv = 1.0 / (theta ** (torch.arange(0, dim, 2)[: dim // 2].double() / dim))
    ang = torch.outer(torch.arange(end), inv)
    return torch.polar(torch.ones_like(ang), ang)


def rope_tables_3d(head_dim: int):
    """wan_video_dit.py:75-80 -- (f, h, w) axis tables with dims head_dim-2*(head_dim//3), head_dim//3, head_dim//3."""
    third = head_dim // 3
    return rope_axis_table(head_dim - 2 * third), rope_axis_table(third), rope_axis_table(third)


def rope_freqs(head_dim: int, f: int, h: int, w: int, device) -> torch.Tensor:
    """wan_video_dit.py:380-384 / model_fn :1474-1478 -- (f*h*w, 1, head_dim/2) complex128."""
    tf, th, tw = rope_tables_3d(head_dim)
    return torch.cat([
        tf[:f].view(f, 1, 1, -1).expand(f, h, w, -1),
        th[:h].view(1, h, 1, -1).expand(f, h, w, -1),
        tw[:w].view(1, 1, w, -1).expand(f, h, w, -1),
    ], dim=-1).reshape(f * h * w, 1, -1).to(device)


def rope_apply(x: torch.Tensor, freqs: torch.Tensor, num_heads: int) -> torch.Tensor:
    """wan_video_dit.py:92-97 -- interleaved pairs as complex128, one rounding back to x.dtype."""
    b, s, _ = x.shape
    xc = torch.view_as_complex(x.to(torch.float64).reshape(b, s, num_heads, -1, 2))
    return torch.view_as_real(xc * freqs).flatten(2).to(x.dtype)


def rms_norm(x: torch.Tensor, weight: torch.Tensor, eps: float) -> torch.Tensor:
    """wan_video_dit.py:100-111 -- fp32 normalisation over the full row, cast back, then * weight."""
    xf = x.float()
    return (xf * torch.rsqrt(xf.pow(2).mean(dim=-1, keepdim=True) + eps)).to(x.dtype) * weight


def attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, num_heads: int) -> torch.Tensor:
    """wan_video_dit.py:55-60 -- b s (n d) -> b n s d, SDPA (no mask, scale 1/sqrt(d)), back."""
    b, sq, d = q.shape
    hd = d // num_heads
    qh = q.view(b, sq, num_heads, hd).transpose(1, 2)
    kh = k.view(b, k.shape[1], num_heads, hd).transpose(1, 2)
    vh = v.view(b, v.shape[1], num_heads, hd).transpose(1, 2)
    o = F.scaled_dot_product_attention(qh, kh, vh)
    return o.transpose(1, 2).reshape(b, sq, d)


def _lin(sd, prefix, x):
    return F.linear(x, sd[prefix + ".weight"], sd[prefix + ".bias"])


def self_attention(sd, pre: str, x, freqs, cfg: DiTConfig):
    """SelfAttention.forward, wan_video_dit.py:140-147."""
    q = rms_norm(_lin(sd, pre + ".q", x), sd[pre + ".norm_q.weight"], cfg.eps)
    k = rms_norm(_lin(sd, pre + ".k", x), sd[pre + ".norm_k.weight"], cfg.eps)
    v = _lin(sd, pre + ".v", x)
    q = rope_apply(q, freqs, cfg.num_heads)
    k = rope_apply(k, freqs, cfg.num_heads)
    return _lin(sd, pre + ".o", attention(q, k, v, cfg.num_heads))


def cross_attention(sd, pre: str, x, ctx, cfg: DiTConfig):
    """CrossAttention.forward without the image branch (has_image_input=False), wan_video_dit.py:171-186."""
    q = rms_norm(_lin(sd, pre + ".q", x), sd[pre + ".norm_q.weight"], cfg.eps)
    k = rms_norm(_lin(sd, pre + ".k", ctx), sd[pre + ".norm_k.weight"], cfg.eps)
    v = _lin(sd, pre + ".v", ctx)
    return _lin(sd, pre + ".o", attention(q, k, v, cfg.num_heads))


def dit_block(sd, pre: str, x, context, t_mod, freqs, cfg: DiTConfig):
    """DiTBlock.forward, wan_video_dit.py:214-230 (t_mod is (B, 6, dim))."""
    mod = sd[pre + ".modulation"].to(dtype=t_mod.dtype, device=t_mod.device) + t_mod
    shift_msa, scale_msa, gate_msa, shift_mlp, scale_mlp, gate_mlp = mod.chunk(6, dim=1)
    d = (cfg.dim,)
    h = F.layer_norm(x, d, eps=cfg.eps) * (1 + scale_msa) + shift_msa
    x = x + gate_msa * self_attention(sd, pre + ".self_attn", h, freqs, cfg)
    n3 = F.layer_norm(x, d, sd[pre + ".norm3.weight"], sd[pre + ".norm3.bias"], eps=cfg.eps)
    x = x + cross_attention(sd, pre + ".cross_attn", n3, context, cfg)
    h = F.layer_norm(x, d, eps=cfg.eps) * (1 + scale_mlp) + shift_mlp
    ff = _lin(sd, pre + ".ffn.2", F.gelu(_lin(sd, pre + ".ffn.0", h), approximate="tanh"))
    return x + gate_mlp * ff


def patchify(weight, bias, x):
    """WanModel.patchify, wan_video_dit.py:341-349: Conv3d k=s=(1,2,2) then 'b c f h w -> b (f h w) c'."""
    y = F.conv3d(x, weight, bias, stride=tuple(weight.shape[2:]))
    grid = tuple(y.shape[2:])
    return y.flatten(2).transpose(1, 2).contiguous(), grid


def head(sd, x, t, cfg: DiTConfig):
    """Head.forward (2-D t branch), wan_video_dit.py:262-269."""
    shift, scale = (sd["head.modulation"].to(dtype=t.dtype, device=t.device) + t).chunk(2, dim=1)
    return _lin(sd, "head.head", F.layer_norm(x, (cfg.dim,), eps=cfg.eps) * (1 + scale) + shift)


def unpatchify(x, grid, cfg: DiTConfig):
    """WanModel.unpatchify, wan_video_dit.py:351-356: 'b (f h w) (x y z c) -> b c (f x) (h y) (w z)'."""
    f, h, w = grid
    px, py, pz = cfg.patch_size
    b = x.shape[0]
    c = x.shape[-1] // (px * py * pz)
    x = x.view(b, f, h, w, px, py, pz, c).permute(0, 7, 1, 4, 2, 5, 3, 6)
    return x.reshape(b, c, f * px, h * py, w * pz)


# ------------------------------------------------------------------------------------------------- the whole forward
def model_fn(sd, cfg: DiTConfig, latents, timestep, context, y=None, controlnet_sd=None,
             control_signal_video_latents=None, controlnet_num_layers: int = 0, controlnet_stride=None,
             return_intermediates: bool = False):
    """model_fn_wan_video (src/goal_force/wan_video_new.py:1349-1591) for the goal-force configuration:
    no motion controller / VACE / TeaCache / USP / clip feature; ControlNet optional.
    """
    t = _time_embedding(sd, cfg, timestep)                                         # :1441
    t_mod = F.linear(F.silu(t), sd["time_projection.1.weight"], sd["time_projection.1.bias"]).unflatten(1, (6, cfg.dim))
    ctx = _lin(sd, "text_embedding.2", F.gelu(_lin(sd, "text_embedding.0", context), approximate="tanh"))  # :1447
    x = latents
    if y is not None:                                                              # :1457-1458
        x = torch.cat([x, y], dim=1)
    x, (f, h, w) = patchify(sd["patch_embedding.weight"], sd["patch_embedding.bias"], x)   # :1464
    freqs = rope_freqs(cfg.head_dim, f, h, w, x.device)                            # :1474-1478
    inter = {}
    states = []
    if controlnet_sd is not None:                                                  # :1489-1522
        s, _ = patchify(controlnet_sd["controlnet_patch_embedding.patch_embedding.weight"],
                        controlnet_sd["controlnet_patch_embedding.patch_embedding.bias"],
                        control_signal_video_latents)
        for i in range(controlnet_num_layers):
            s = dit_block(controlnet_sd, f"controlnet_dit.blocks.{i}", s, ctx, t_mod, freqs, cfg)
            states.append(s)
    for i in range(cfg.num_layers):                                                # :1540-1570
        x = dit_block(sd, f"blocks.{i}", x, ctx, t_mod, freqs, cfg)
        if controlnet_sd is not None:
            if controlnet_stride is not None:
                if i % controlnet_stride == 0 and i // controlnet_stride < len(states):
                    x = x + states[i // controlnet_stride]
            elif i < controlnet_num_layers:
                zw = controlnet_sd[f"controlnet_zero_convs_after.{i}.weight"]
                zb = controlnet_sd[f"controlnet_zero_convs_after.{i}.bias"]
                x = x + F.conv1d(states[i].transpose(1, 2), zw, zb).transpose(1, 2)
        if return_intermediates:
            inter[f"block{i}"] = x
    x = head(sd, x, t, cfg)                                                        # :1581
    out = unpatchify(x, (f, h, w), cfg)                                            # :1590
    return (out, inter) if return_intermediates else out


def _time_embedding(sd, cfg, timestep):
    e = sinusoidal_embedding_1d(cfg.freq_dim, timestep)
    return _lin(sd, "time_embedding.2", F.silu(_lin(sd, "time_embedding.0", e)))


# --------------------------------------------------------------------------------------------- seeded random weights
def random_state_dict(cfg: DiTConfig, seed: int = 0, dtype=torch.float32, device="cpu") -> dict:
    """Deterministic random-init DiT weights with the reference's key names (wan_video_dit.py:307-326).
    Not the nn.Module default init (that needs the reference classes); scaled so activations stay O(1)."""
    g = torch.Generator("cpu").manual_seed(seed)

    def rn(*shape, scale=1.0):
        return (torch.randn(*shape, generator=g) * scale)

    sd = {}
    k_in = cfg.in_dim * 4
    sd["patch_embedding.weight"] = rn(cfg.dim, cfg.in_dim, 1, 2, 2, scale=k_in ** -0.5)
    sd["patch_embedding.bias"] = rn(cfg.dim, scale=0.02)

    def linear(name, out_f, in_f):
        sd[name + ".weight"] = rn(out_f, in_f, scale=in_f ** -0.5)
        sd[name + ".bias"] = rn(out_f, scale=0.02)

    linear("text_embedding.0", cfg.dim, cfg.text_dim)
    linear("text_embedding.2", cfg.dim, cfg.dim)
    linear("time_embedding.0", cfg.dim, cfg.freq_dim)
    linear("time_embedding.2", cfg.dim, cfg.dim)
    linear("time_projection.1", cfg.dim * 6, cfg.dim)
    for i in range(cfg.num_layers):
        _random_block(sd, f"blocks.{i}", cfg, rn, linear)
    linear("head.head", cfg.out_dim * 4, cfg.dim)
    sd["head.modulation"] = rn(1, 2, cfg.dim, scale=cfg.dim ** -0.5)
    return {k: v.to(dtype=dtype, device=device) for k, v in sd.items()}


def _random_block(sd, pre, cfg, rn, linear):
    for att in ("self_attn", "cross_attn"):
        for p in "qkvo":
            linear(f"{pre}.{att}.{p}", cfg.dim, cfg.dim)
        sd[f"{pre}.{att}.norm_q.weight"] = 1.0 + rn(cfg.dim, scale=0.1)
        sd[f"{pre}.{att}.norm_k.weight"] = 1.0 + rn(cfg.dim, scale=0.1)
    sd[f"{pre}.norm3.weight"] = 1.0 + rn(cfg.dim, scale=0.1)
    sd[f"{pre}.norm3.bias"] = rn(cfg.dim, scale=0.05)
    linear(f"{pre}.ffn.0", cfg.ffn_dim, cfg.dim)
    linear(f"{pre}.ffn.2", cfg.dim, cfg.ffn_dim)
    sd[f"{pre}.modulation"] = rn(1, 6, cfg.dim, scale=cfg.dim ** -0.5)


def random_controlnet_state_dict(cfg: DiTConfig, num_layers: int, seed: int = 1, dtype=torch.float32, device="cpu",
                                 zero_convs: bool = False, control_in_dim: int = 16) -> dict:
    """ControlNet weights with the reference's key names (src/goal_force/wan_video_new.py:97-117). The zero-convs are
    given small non-zero weights unless zero_convs=True, otherwise the branch would be numerically invisible."""
    g = torch.Generator("cpu").manual_seed(seed)

    def rn(*shape, scale=1.0):
        return torch.randn(*shape, generator=g) * scale

    sd = {}

    def linear(name, out_f, in_f):
        sd[name + ".weight"] = rn(out_f, in_f, scale=in_f ** -0.5)
        sd[name + ".bias"] = rn(out_f, scale=0.02)

    sd["controlnet_patch_embedding.patch_embedding.weight"] = rn(cfg.dim, control_in_dim, 1, 2, 2,
                                                                 scale=(control_in_dim * 4) ** -0.5)
    sd["controlnet_patch_embedding.patch_embedding.bias"] = rn(cfg.dim, scale=0.02)
    for i in range(num_layers):
        _random_block(sd, f"controlnet_dit.blocks.{i}", cfg, rn, linear)
        if zero_convs:
            sd[f"controlnet_zero_convs_after.{i}.weight"] = torch.zeros(cfg.dim, cfg.dim, 1)
            sd[f"controlnet_zero_convs_after.{i}.bias"] = torch.zeros(cfg.dim)
        else:
            sd[f"controlnet_zero_convs_after.{i}.weight"] = rn(cfg.dim, cfg.dim, 1, scale=0.5 * cfg.dim ** -0.5)
            sd[f"controlnet_zero_convs_after.{i}.bias"] = rn(cfg.dim, scale=0.02)
    return {k: v.to(dtype=dtype, device=device) for k, v in sd.items()}


def synthetic_inputs(cfg: DiTConfig, frames_lat: int, h_lat: int, w_lat: int, seed: int = 1, dtype=torch.float32,
                     device="cpu", timestep: float = 900.0, with_y: bool | None = None, ctx_len: int = 512,
                     ctx_valid: int = 64) -> dict:
    """SURVEY 8(d) synthetic inputs: latents ~ N(0,1); y = 4 mask channels (1 at latent frame 0) + N(0,1) image latents;
    context ~ N(0,1) with rows >= ctx_valid zeroed (wan_prompter.py:105-108); timestep rounded to bf16 as the pipeline
    does (src/goal_force/wan_video_new.py:707)."""
    g = torch.Generator("cpu").manual_seed(seed)
    with_y = (cfg.in_dim > 16) if with_y is None else with_y
    out = {"latents": torch.randn(1, 16, frames_lat, h_lat, w_lat, generator=g)}
    if with_y:
        msk = torch.zeros(1, 4, frames_lat, h_lat, w_lat)
        msk[:, :, 0] = 1.0
        out["y"] = torch.cat([msk, torch.randn(1, cfg.in_dim - 20, frames_lat, h_lat, w_lat, generator=g)], dim=1)
    ctx = torch.randn(1, ctx_len, cfg.text_dim, generator=g)
    ctx[:, ctx_valid:] = 0
    out["context"] = ctx
    out["control_signal_video_latents"] = torch.randn(1, 16, frames_lat, h_lat, w_lat, generator=g)
    out = {k: v.to(dtype=dtype, device=device) for k, v in out.items()}
    out["timestep"] = torch.tensor([timestep], dtype=torch.bfloat16).to(dtype=dtype, device=device)
    return out


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.double().flatten().cpu(), b.double().flatten().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def cosine(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.double().flatten().cpu(), b.double().flatten().cpu()
    return float(torch.dot(a, b) / (a.norm() * b.norm()).clamp_min(1e-30))
