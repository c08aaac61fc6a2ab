"""Flow-matching scheduler mirror of diffsynth/schedulers/flow_match.py:5-82 (inference subset).

Host logic only: the sigma table is a handful of fp32 scalars built exactly as the reference builds them
(torch.linspace on CPU, shifted sigma = s*x / (1 + (s-1)*x)); the per-step latent update itself runs on the GPU
in gf_cfg_euler_bf16 (goal_force_b200.capi.cfg_euler).
"""
from __future__ import annotations

import torch


class FlowMatchScheduler:
    """Same constructor defaults the goal-force pipeline uses (src/goal_force/wan_video_new.py:129):
    FlowMatchScheduler(shift=5, sigma_min=0.0, extra_one_step=True)."""

    def __init__(self, num_inference_steps: int = 100, num_train_timesteps: int = 1000, shift: float = 5.0,
                 sigma_max: float = 1.0, sigma_min: float = 0.0, extra_one_step: bool = True):
        self.num_train_timesteps = num_train_timesteps
        self.shift = shift
        self.sigma_max = sigma_max
        self.sigma_min = sigma_min
        self.extra_one_step = extra_one_step
        self.set_timesteps(num_inference_steps)

    def set_timesteps(self, num_inference_steps: int = 100, denoising_strength: float = 1.0, shift: float | None = None):
        """flow_match.py:34-60 without the training / exponential / terminal-shift branches (unused at inference)."""
        if shift is not None:
            self.shift = shift
        start = self.sigma_min + (self.sigma_max - self.sigma_min) * denoising_strength
        if self.extra_one_step:
            sig = torch.linspace(start, self.sigma_min, num_inference_steps + 1)[:-1]
        else:
            sig = torch.linspace(start, self.sigma_min, num_inference_steps)
        self.sigmas = self.shift * sig / (1 + (self.shift - 1) * sig)
        self.timesteps = self.sigmas * self.num_train_timesteps

    def timestep_index(self, timestep) -> int:
        if isinstance(timestep, torch.Tensor):
            timestep = timestep.detach().float().cpu()
        return int(torch.argmin((self.timesteps - timestep).abs()))

    def sigma_pair(self, timestep, to_final: bool = False):
        """(sigma, sigma_next) as the reference's step() picks them (flow_match.py:72-80)."""
        i = self.timestep_index(timestep)
        sigma = self.sigmas[i]
        if to_final or i + 1 >= len(self.timesteps):
            nxt = torch.zeros(())
        else:
            nxt = self.sigmas[i + 1]
        return sigma, nxt

    def dsigma(self, timestep, to_final: bool = False) -> float:
        """fp32 value of (sigma_next - sigma), the scalar the reference multiplies model_output by."""
        s, n = self.sigma_pair(timestep, to_final)
        return float((n - s).to(torch.float32))

    def step(self, model_output: torch.Tensor, timestep, sample: torch.Tensor, to_final: bool = False, **kwargs):
        """prev = sample + model_output * (sigma_next - sigma) on the GPU (flow_match.py:81), no CFG."""
        from . import capi
        return capi.cfg_euler(model_output.contiguous(), None, sample.contiguous(), 1.0, self.dsigma(timestep, to_final))
