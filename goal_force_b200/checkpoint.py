"""Checkpoint ingestion without DiffSynth's model manager: safetensors shards -> the key/shape contract that
WanModelB200 / ControlNetB200 consume (the reference's own state_dict names).

Mirrors what the reference does around loading, nothing else:
  * `load_state_dict(path)` + strip of the 'pipe.controlnet.' prefix in load_controlnet_weights
    (src/goal_force/wan_video_new.py:176-178) -- the prefix is accepted as-is by ControlNetB200;
  * a DiT expert stored as several *.safetensors shards (ModelConfig(path=[...]),
    scripts/inference/inference_goal_force.py:83-97) is presented as one mapping.
Tensors are read lazily, one key at a time (`safe_open`), so a 28 GB expert never sits in host memory twice: the model
constructors pull each weight, re-lay it out on the GPU and drop the source tensor.
"""
from __future__ import annotations

from pathlib import Path
from typing import Iterable

import torch


class SafetensorsStateDict:
    """Read-only mapping key -> tensor over one or more .safetensors files (first file that holds a key wins)."""

    def __init__(self, paths: str | Path | Iterable[str | Path], device: str = "cpu"):
        from safetensors import safe_open
        if isinstance(paths, (str, Path)):
            paths = [paths]
        self.paths = [Path(p) for p in paths]
        if not self.paths:
            raise ValueError("no checkpoint files given")
        self.device = device
        self._open = safe_open
        self._where: dict = {}
        for p in self.paths:
            if not p.exists():
                raise FileNotFoundError(p)
            with safe_open(str(p), framework="pt", device="cpu") as f:
                for k in f.keys():
                    self._where.setdefault(k, p)

    def __contains__(self, key: str) -> bool:
        return key in self._where

    def __len__(self) -> int:
        return len(self._where)

    def keys(self):
        return self._where.keys()

    def __getitem__(self, key: str) -> torch.Tensor:
        p = self._where.get(key)
        if p is None:
            raise KeyError(key)
        with self._open(str(p), framework="pt", device=self.device) as f:
            return f.get_tensor(key)


def load_dit(paths, cfg, device="cuda"):
    """WanModelB200 from the safetensors shards of one expert (keys as in WanModel.state_dict())."""
    from .wan_dit import WanModelB200
    return WanModelB200(cfg, SafetensorsStateDict(paths), device=device)


def load_controlnet(path, cfg, num_layers: int, stride=None, device="cuda"):
    """ControlNetB200 from a goal-force training checkpoint (step-N.safetensors; keys carry 'pipe.controlnet.')."""
    from .wan_dit import ControlNetB200
    return ControlNetB200(cfg, SafetensorsStateDict(path), num_layers, stride=stride, device=device)


def read_vae_state_dict(path) -> dict:
    """VideoVAE_ state dict from a Wan2.1 VAE checkpoint, in the key names WanVideoVAEB200 consumes.  Accepts what the
    reference's loader accepts (diffsynth/models/wan_video_vae.py:1256-1267): a torch pickle (Wan2.1_VAE.pth), optionally
    wrapped in {'model_state': ...}, or a .safetensors file; keys with the wrapper's 'model.' prefix (what
    WanVideoVAEStateDictConverter.from_civitai produces for WanVideoVAE) are accepted as well.  254 MB: read eagerly."""
    path = Path(path)
    if not path.exists():
        raise FileNotFoundError(path)
    if path.suffix == ".safetensors":
        from safetensors.torch import load_file
        sd = load_file(str(path), device="cpu")
    else:
        sd = torch.load(str(path), map_location="cpu", weights_only=True)
    if "model_state" in sd:
        sd = sd["model_state"]
    if sd and all(k.startswith("model.") for k in sd):
        sd = {k[len("model."):]: v for k, v in sd.items()}
    for need in ("encoder.conv1.weight", "decoder.conv1.weight", "conv1.weight", "conv2.weight", "decoder.head.2.weight"):
        if need not in sd:
            raise KeyError(f"{path}: not a Wan2.1 VideoVAE_ checkpoint (missing {need})")
    return sd


def load_vae(path, device="cuda"):
    """WanVideoVAEB200 from a Wan2.1 VAE checkpoint file (width and z_dim are read off the weights)."""
    from .wan_vae import WanVideoVAEB200
    sd = read_vae_state_dict(path)
    return WanVideoVAEB200(sd, dim=sd["encoder.conv1.weight"].shape[0], z_dim=sd["conv2.weight"].shape[0], device=device)
