"""Synthetic (random-init) weights and inputs of the reference's shapes, generated directly on the GPU.

BASELINE.json prescribes random-init weights of the A14B architecture and synthetic latents (there is no network
for checkpoints). A14B is 14 B parameters per expert, so the weights are produced lazily, key by key, on the device:
`LazyRandomStateDict` looks like the reference's state_dict (same key names and shapes, wan_video_dit.py:307-326 and
src/goal_force/wan_video_new.py:97-117) to WanModelB200 / ControlNetB200 but never holds more than one tensor.
"""
from __future__ import annotations

import re
import zlib

import torch

from .wan_dit import DiTConfig


class LazyRandomStateDict:
    """Mapping key -> freshly generated bf16 tensor. Scales follow nn.Linear-like fan-in so activations stay O(1)."""

    def __init__(self, cfg: DiTConfig, seed: int = 0, device="cuda", controlnet_layers: int = 0,
                 zero_convs: bool = False, control_in_dim: int = 16):
        self.cfg, self.seed, self.device = cfg, seed, torch.device(device)
        self.controlnet_layers, self.zero_convs, self.control_in_dim = controlnet_layers, zero_convs, control_in_dim

    def _shape_scale(self, key: str):
        c = self.cfg
        d = c.dim
        key = key.replace("pipe.controlnet.", "", 1)
        if key == "controlnet_patch_embedding.patch_embedding.weight":
            return (d, self.control_in_dim, 1, 2, 2), (self.control_in_dim * 4) ** -0.5, 0.0
        if key == "controlnet_patch_embedding.patch_embedding.bias":
            return (d,), 0.02, 0.0
        m = re.match(r"controlnet_zero_convs_after\.\d+\.(weight|bias)", key)
        if m:
            if m.group(1) == "weight":
                return (d, d, 1), 0.0 if self.zero_convs else 0.5 * d ** -0.5, 0.0
            return (d,), 0.0 if self.zero_convs else 0.02, 0.0
        key = re.sub(r"^controlnet_dit\.", "", key)
        table = {
            "patch_embedding.weight": ((d, c.in_dim, 1, 2, 2), (c.in_dim * 4) ** -0.5, 0.0),
            "patch_embedding.bias": ((d,), 0.02, 0.0),
            "text_embedding.0.weight": ((d, c.text_dim), c.text_dim ** -0.5, 0.0),
            "text_embedding.0.bias": ((d,), 0.02, 0.0),
            "text_embedding.2.weight": ((d, d), d ** -0.5, 0.0),
            "text_embedding.2.bias": ((d,), 0.02, 0.0),
            "time_embedding.0.weight": ((d, c.freq_dim), c.freq_dim ** -0.5, 0.0),
            "time_embedding.0.bias": ((d,), 0.02, 0.0),
            "time_embedding.2.weight": ((d, d), d ** -0.5, 0.0),
            "time_embedding.2.bias": ((d,), 0.02, 0.0),
            "time_projection.1.weight": ((6 * d, d), d ** -0.5, 0.0),
            "time_projection.1.bias": ((6 * d,), 0.02, 0.0),
            "head.head.weight": ((c.out_dim * 4, d), d ** -0.5, 0.0),
            "head.head.bias": ((c.out_dim * 4,), 0.02, 0.0),
            "head.modulation": ((1, 2, d), d ** -0.5, 0.0),
        }
        if key in table:
            return table[key]
        m = re.match(r"blocks\.\d+\.(.+)", key)
        if not m:
            raise KeyError(key)
        sub = m.group(1)
        if re.match(r"(self_attn|cross_attn)\.[qkvo]\.weight", sub):
            return (d, d), d ** -0.5, 0.0
        if re.match(r"(self_attn|cross_attn)\.[qkvo]\.bias", sub):
            return (d,), 0.02, 0.0
        if re.match(r"(self_attn|cross_attn)\.norm_[qk]\.weight", sub) or sub == "norm3.weight":
            return (d,), 0.1, 1.0
        blk = {"norm3.bias": ((d,), 0.05, 0.0), "ffn.0.weight": ((c.ffn_dim, d), d ** -0.5, 0.0),
               "ffn.0.bias": ((c.ffn_dim,), 0.02, 0.0), "ffn.2.weight": ((d, c.ffn_dim), c.ffn_dim ** -0.5, 0.0),
               "ffn.2.bias": ((d,), 0.02, 0.0), "modulation": ((1, 6, d), d ** -0.5, 0.0)}
        if sub in blk:
            return blk[sub]
        raise KeyError(key)

    def __getitem__(self, key: str) -> torch.Tensor:
        shape, scale, offset = self._shape_scale(key)
        g = torch.Generator(self.device).manual_seed((self.seed * 1000003 + zlib.crc32(key.encode())) & 0x7FFFFFFF)
        t = torch.randn(shape, generator=g, device=self.device, dtype=torch.bfloat16)
        if scale != 1.0:
            t.mul_(scale)
        if offset:
            t.add_(offset)
        return t

    def items(self):
        raise TypeError("LazyRandomStateDict is index-only (it never materialises the full model)")


def synthetic_inputs(cfg: DiTConfig, frames_lat: int, h_lat: int, w_lat: int, seed: int = 1, device="cpu",
                     ctx_len: int = 512, ctx_valid: int = 64, timestep: float = 900.0, pin: bool = False) -> dict:
    """SURVEY 8(d) synthetic step inputs (bf16): latents, y (4 mask + 16 image-latent channels for I2V), context
    (rows >= ctx_valid zeroed, as wan_prompter.py:105-108 does), control latents, bf16 timestep."""
    g = torch.Generator("cpu").manual_seed(seed)
    out = {"latents": torch.randn(1, 16, frames_lat, h_lat, w_lat, generator=g)}
    if cfg.in_dim > 16:
        msk = torch.zeros(1, 4, frames_lat, h_lat, w_lat)
        msk[:, :, 0] = 1.0
        out["y"] = torch.cat([msk, torch.randn(1, cfg.in_dim - 20, frames_lat, h_lat, w_lat, generator=g)], dim=1)
    ctx = torch.randn(1, ctx_len, cfg.text_dim, generator=g)
    ctx[:, ctx_valid:] = 0
    out["context"] = ctx
    out["control_signal_video_latents"] = torch.randn(1, 16, frames_lat, h_lat, w_lat, generator=g)
    out["timestep"] = torch.tensor([timestep])
    out = {k: v.to(torch.bfloat16) for k, v in out.items()}
    if pin and torch.cuda.is_available():
        out = {k: v.pin_memory() for k, v in out.items()}
    if str(device) != "cpu":
        out = {k: v.to(device) for k, v in out.items()}
    return out
