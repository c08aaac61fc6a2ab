"""nn.Module stand-ins with the reference's attribute tree and state_dict key names -- TEST INFRASTRUCTURE ONLY (only tests/ import this).

The GPU box has no /root/reference, so the drop-in recipe of INTEGRATION.md (`pipe.model_fn = model_fn_wan_video`
with the pipeline's own `pipe.dit` / `pipe.controlnet` nn.Modules) is exercised there with these classes. They carry
exactly what goal_force_b200 reads from a reference module (SURVEY 8b attribute surface):
  WanModel  (diffsynth/models/wan_video_dit.py:270-339): dim, in_dim, freq_dim, patch_size, has_image_input,
            seperated_timestep, require_vae_embedding, require_clip_embedding, fuse_vae_embedding_in_latents,
            has_image_pos_emb, has_ref_conv, control_adapter, patch_embedding (Conv3d), text_embedding / time_embedding /
            time_projection (Sequential), blocks[i] = DiTBlock(dim, num_heads, ffn_dim, norm1.eps, self_attn / cross_attn
            with q,k,v,o Linear + norm_q/norm_k RMSNorm, norm3, ffn, modulation), head.head, head.modulation
  ControlNet (src/goal_force/wan_video_new.py:97-117): num_layers, stride, controlnet_patch_embedding.patch_embedding,
            controlnet_dit.blocks, controlnet_zero_convs_after (Conv1d k=1)
tests/test_host_logic_cpu.py::test_standins_mirror_live_reference pins them against the live reference classes
(same state_dict keys and shapes, same attributes) whenever /root/reference is present. They have no forward():
the product never calls a reference module, it only reads weights and attributes.
"""
from __future__ import annotations

import torch
from torch import nn


class _RMSNorm(nn.Module):
    def __init__(self, dim, eps):
        super().__init__()
        self.eps = eps
        self.weight = nn.Parameter(torch.ones(dim))


class _Attention(nn.Module):
    def __init__(self, dim, num_heads, eps):
        super().__init__()
        self.dim, self.num_heads, self.head_dim = dim, num_heads, dim // num_heads
        self.q, self.k, self.v, self.o = (nn.Linear(dim, dim) for _ in range(4))
        self.norm_q, self.norm_k = _RMSNorm(dim, eps), _RMSNorm(dim, eps)


class DiTBlockStandIn(nn.Module):
    def __init__(self, has_image_input, dim, num_heads, ffn_dim, eps):
        super().__init__()
        self.dim, self.num_heads, self.ffn_dim = dim, num_heads, ffn_dim
        self.self_attn = _Attention(dim, num_heads, eps)
        self.cross_attn = _Attention(dim, num_heads, eps)
        self.norm1 = nn.LayerNorm(dim, eps=eps, elementwise_affine=False)
        self.norm2 = nn.LayerNorm(dim, eps=eps, elementwise_affine=False)
        self.norm3 = nn.LayerNorm(dim, eps=eps)
        self.ffn = nn.Sequential(nn.Linear(dim, ffn_dim), nn.GELU(approximate="tanh"), nn.Linear(ffn_dim, dim))
        self.modulation = nn.Parameter(torch.randn(1, 6, dim) / dim ** 0.5)


class _Head(nn.Module):
    def __init__(self, dim, out_dim, patch_size, eps):
        super().__init__()
        self.norm = nn.LayerNorm(dim, eps=eps, elementwise_affine=False)
        self.head = nn.Linear(dim, out_dim * patch_size[0] * patch_size[1] * patch_size[2])
        self.modulation = nn.Parameter(torch.randn(1, 2, dim) / dim ** 0.5)


class WanModelStandIn(nn.Module):
    def __init__(self, dim, in_dim, ffn_dim, out_dim, text_dim, freq_dim, eps, patch_size, num_heads, num_layers,
                 has_image_input=False, require_vae_embedding=True, require_clip_embedding=False):
        super().__init__()
        self.dim, self.in_dim, self.freq_dim, self.patch_size = dim, in_dim, freq_dim, tuple(patch_size)
        self.has_image_input = has_image_input
        self.seperated_timestep = False
        self.require_vae_embedding = require_vae_embedding
        self.require_clip_embedding = require_clip_embedding
        self.fuse_vae_embedding_in_latents = False
        self.has_image_pos_emb = False
        self.has_ref_conv = False
        self.control_adapter = None
        self.patch_embedding = nn.Conv3d(in_dim, dim, kernel_size=tuple(patch_size), stride=tuple(patch_size))
        self.text_embedding = nn.Sequential(nn.Linear(text_dim, dim), nn.GELU(approximate="tanh"), nn.Linear(dim, dim))
        self.time_embedding = nn.Sequential(nn.Linear(freq_dim, dim), nn.SiLU(), nn.Linear(dim, dim))
        self.time_projection = nn.Sequential(nn.SiLU(), nn.Linear(dim, dim * 6))
        self.blocks = nn.ModuleList([DiTBlockStandIn(has_image_input, dim, num_heads, ffn_dim, eps)
                                     for _ in range(num_layers)])
        self.head = _Head(dim, out_dim, tuple(patch_size), eps)


class _CNPatch(nn.Module):
    def __init__(self, in_channels, dim, patch_size):
        super().__init__()
        self.patch_embedding = nn.Conv3d(in_channels, dim, kernel_size=patch_size, stride=patch_size)


class _CNDiT(nn.Module):
    def __init__(self, num_layers, dim, num_heads, ffn_dim, eps):
        super().__init__()
        self.num_layers = num_layers
        self.blocks = nn.ModuleList([DiTBlockStandIn(False, dim, num_heads, ffn_dim, eps) for _ in range(num_layers)])


class ControlNetStandIn(nn.Module):
    """The reference hard-codes dim 5120 / 40 heads / ffn 13824 (wan_video_new.py:56-60); the stand-in takes them as
    arguments so that small test shapes are possible, with the reference's values as defaults."""

    def __init__(self, num_layers, stride=None, torch_dtype=torch.bfloat16, dim=5120, num_heads=40, ffn_dim=13824,
                 eps=1e-6):
        super().__init__()
        self.num_layers, self.stride = num_layers, stride
        self.controlnet_patch_embedding = _CNPatch(16, dim, (1, 2, 2)).to(torch_dtype)
        self.controlnet_dit = _CNDiT(num_layers, dim, num_heads, ffn_dim, eps)
        self.controlnet_zero_convs_after = nn.ModuleList([nn.Conv1d(dim, dim, kernel_size=1, dtype=torch_dtype)
                                                          for _ in range(num_layers)])
        for m in self.controlnet_zero_convs_after:          # zero_module (wan_video_new.py:40-46)
            for p in m.parameters():
                p.detach().zero_()


def wan_standin_from_cfg(cfg, state_dict=None, dtype=torch.bfloat16, device="cpu") -> WanModelStandIn:
    m = WanModelStandIn(dim=cfg.dim, in_dim=cfg.in_dim, ffn_dim=cfg.ffn_dim, out_dim=cfg.out_dim, text_dim=cfg.text_dim,
                        freq_dim=cfg.freq_dim, eps=cfg.eps, patch_size=cfg.patch_size, num_heads=cfg.num_heads,
                        num_layers=cfg.num_layers)
    if state_dict is not None:
        m.load_state_dict(state_dict, strict=True)
    return m.to(device=device, dtype=dtype).eval()


def controlnet_standin_from_cfg(cfg, num_layers, state_dict=None, stride=None, dtype=torch.bfloat16,
                                device="cpu") -> ControlNetStandIn:
    m = ControlNetStandIn(num_layers, stride=stride, torch_dtype=dtype, dim=cfg.dim, num_heads=cfg.num_heads,
                          ffn_dim=cfg.ffn_dim, eps=cfg.eps)
    if state_dict is not None:
        m.load_state_dict(state_dict, strict=True)
    return m.to(device=device, dtype=dtype).eval()


class _ParamTree(nn.Module):
    """A module tree built from dotted parameter names (numeric components become attribute names of sub-modules, as
    nn.Sequential / nn.ModuleList register them), so that state_dict() returns exactly the given keys."""

    def __init__(self, state_dict: dict):
        super().__init__()
        for name, value in state_dict.items():
            mod = self
            parts = name.split(".")
            for part in parts[:-1]:
                if part not in mod._modules:
                    mod.add_module(part, _ParamTree({}))
                mod = mod._modules[part]
            mod.register_parameter(parts[-1], nn.Parameter(value.clone(), requires_grad=False))


class VideoVAEStandIn(_ParamTree):
    """VideoVAE_ (diffsynth/models/wan_video_vae.py:842-1055) as goal_force_b200.wan_vae reads it: `.z_dim`,
    `.encoder.conv1.weight` (width) and `state_dict()` with the reference's key names."""

    def __init__(self, state_dict: dict, z_dim: int = 16):
        super().__init__(state_dict)
        self.z_dim = z_dim


class WanVideoVAEStandIn(nn.Module):
    """WanVideoVAE (:1057-1080): `.model` (VideoVAE_), `.upsampling_factor`, `.z_dim`."""

    def __init__(self, state_dict: dict, z_dim: int = 16):
        super().__init__()
        self.model = VideoVAEStandIn(state_dict, z_dim)
        self.upsampling_factor = 8
        self.z_dim = z_dim
