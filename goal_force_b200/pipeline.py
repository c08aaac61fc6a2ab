"""Denoising loop of the goal-force WanVideoPipeline on B200 kernels.

Mirrors src/goal_force/wan_video_new.py:598-723 (`WanVideoPipeline.__call__`, denoise part) and the conditioning
units that feed it: NoiseInitializer (:751-763), ImageEmbedderVAE's mask construction (:899-910).  VAE / T5 / CLIP
stay outside (north-star scope): the caller passes latents-space tensors, exactly what the reference's units hand to
`model_fn`.

Two parallel axes on one 8xB200 box, both optional (one process per GPU, torch.distributed / NCCL):
  * sequence parallel (Ulysses) inside one forward  -> goal_force_b200.wan_dit.SequenceParallel
  * CFG parallel: the conditional and unconditional forwards of a step run on two groups of GPUs and exchange the
    4 MB noise prediction once per step (absent in the reference, which runs them back to back, :710-716).
"""
from __future__ import annotations

from dataclasses import dataclass

import torch

from . import capi
from .scheduler import FlowMatchScheduler
from .wan_dit import ControlNetB200, SequenceParallel, WanModelB200, model_fn_wan_video


def generate_noise(shape, seed=None, rand_device="cpu", dtype=torch.bfloat16, device="cuda") -> torch.Tensor:
    """BasePipeline.generate_noise (diffsynth/utils/__init__.py:117-122): fp32 randn from a seeded generator on
    `rand_device`, then cast / moved.  Kept in torch so that a seed reproduces the reference's noise bit for bit."""
    gen = None if seed is None else torch.Generator(rand_device).manual_seed(seed)
    noise = torch.randn(shape, generator=gen, device=rand_device, dtype=torch.float32)
    return noise.to(dtype=dtype, device=device)


def latent_shape(num_frames: int, height: int, width: int, z_dim: int = 16, upsampling: int = 8):
    """NoiseInitializer shape rule (wan_video_new.py:756-759)."""
    return (1, z_dim, (num_frames - 1) // 4 + 1, height // upsampling, width // upsampling)


def first_frame_mask(num_frames: int, h8: int, w8: int, end_image: bool = False, device="cpu") -> torch.Tensor:
    """ImageEmbedderVAE mask (wan_video_new.py:899-910): ones at frame 0 (and the last frame with an end image),
    first frame repeated x4, folded to (4, (num_frames+3)/4, h8, w8). Pure index bookkeeping, bit-exact."""
    msk = torch.ones(1, num_frames, h8, w8, device=device)
    msk[:, 1:] = 0
    if end_image:
        msk[:, -1:] = 1
    msk = torch.concat([torch.repeat_interleave(msk[:, 0:1], repeats=4, dim=1), msk[:, 1:]], dim=1)
    msk = msk.view(1, msk.shape[1] // 4, 4, h8, w8)
    return msk.transpose(1, 2)[0]


def image_condition(vae_latents: torch.Tensor, num_frames: int, end_image: bool = False) -> torch.Tensor:
    """y = cat(mask, vae_latents)[None] (wan_video_new.py:912-916): (1, 4+16, T, h8, w8)."""
    _, T, h8, w8 = vae_latents.shape
    msk = first_frame_mask(num_frames, h8, w8, end_image, device=vae_latents.device).to(vae_latents.dtype)
    if msk.shape[1] != T:
        raise ValueError(f"mask has {msk.shape[1]} latent frames, vae_latents {T}")
    return torch.concat([msk, vae_latents]).unsqueeze(0)


@dataclass
class ParallelLayout:
    """rank = replica * (cfg_size * sp_size) + cfg_index * sp_size + sp_index.
    cfg_size in {1, 2}; sp_size must divide the head count and the token count; `replicas` independent copies of the
    (cfg x sp) grid work on different jobs (the reference's only multi-GPU mode: one process per CSV shard,
    scripts/inference/utils.py:25-57)."""
    world_size: int = 1
    rank: int = 0
    cfg_size: int = 1
    replicas: int = 1

    def __post_init__(self):
        if self.cfg_size not in (1, 2) or self.replicas < 1 or self.world_size % (self.cfg_size * self.replicas):
            raise ValueError(f"cfg_size {self.cfg_size} x replicas {self.replicas} does not fit world size "
                             f"{self.world_size}")

    @property
    def group_size(self) -> int:
        """ranks that cooperate on one job"""
        return self.world_size // self.replicas

    @property
    def sp_size(self) -> int:
        return self.group_size // self.cfg_size

    @property
    def replica(self) -> int:
        return self.rank // self.group_size

    @property
    def cfg_index(self) -> int:
        return (self.rank % self.group_size) // self.sp_size

    @property
    def sp_index(self) -> int:
        return self.rank % self.sp_size

    def sp_ranks(self, cfg_index: int | None = None, replica: int | None = None):
        c = self.cfg_index if cfg_index is None else cfg_index
        r = self.replica if replica is None else replica
        base = r * self.group_size + c * self.sp_size
        return [base + i for i in range(self.sp_size)]

    def cfg_ranks(self, sp_index: int | None = None, replica: int | None = None):
        s = self.sp_index if sp_index is None else sp_index
        r = self.replica if replica is None else replica
        return [r * self.group_size + c * self.sp_size + s for c in range(self.cfg_size)]


class ParallelContext:
    """torch.distributed groups for a ParallelLayout (every rank must construct it: new_group is collective and
    every rank creates every group, in the same order)."""

    def __init__(self, layout: ParallelLayout, transport: str = "peer"):
        import torch.distributed as dist
        self.dist = dist
        self.layout = layout
        self.sp_group = None
        self.cfg_group = None
        if layout.world_size > 1:
            for r in range(layout.replicas):
                for c in range(layout.cfg_size):
                    g = dist.new_group(layout.sp_ranks(c, r)) if layout.sp_size > 1 else None
                    if r == layout.replica and c == layout.cfg_index:
                        self.sp_group = g
                for s in range(layout.sp_size):
                    g = dist.new_group(layout.cfg_ranks(s, r)) if layout.cfg_size > 1 else None
                    if r == layout.replica and s == layout.sp_index:
                        self.cfg_group = g
        self.sp = SequenceParallel(self.sp_group, transport=transport) if layout.sp_size > 1 else None


class GoalForceDenoiser:
    """The hot loop: for t in timesteps: model_fn(posi); model_fn(nega); CFG; Euler step; expert switch."""

    def __init__(self, dit: WanModelB200, dit2: WanModelB200 | None = None, controlnet: ControlNetB200 | None = None,
                 controlnet2: ControlNetB200 | None = None, scheduler: FlowMatchScheduler | None = None,
                 parallel: ParallelContext | None = None, model_fn=model_fn_wan_video):
        self.dit, self.dit2 = dit, dit2
        self.controlnet, self.controlnet2 = controlnet, controlnet2
        self.scheduler = scheduler or FlowMatchScheduler(shift=5, sigma_min=0.0, extra_one_step=True)
        self.parallel = parallel
        self.model_fn = model_fn

    def experts_for(self, timestep: float, switch_DiT_boundary: float):
        """wan_video_new.py:699-704: below boundary*1000 the low-noise expert (and its ControlNet) takes over."""
        if self.dit2 is not None and timestep < switch_DiT_boundary * self.scheduler.num_train_timesteps:
            return self.dit2, self.controlnet2
        return self.dit, self.controlnet

    @torch.no_grad()
    def step(self, latents, timestep, context_posi, context_nega, y, control_latents, cfg_scale,
             switch_DiT_boundary=0.875, out=None):
        """One denoising step -> new latents (same shape/dtype); `timestep` is the scheduler's fp32 scalar tensor."""
        dit, cn = self.experts_for(float(timestep), switch_DiT_boundary)
        ts = timestep.reshape(1).to(dtype=torch.bfloat16, device=latents.device)       # :707 (bf16 rounding, F11)
        par = self.parallel
        sp = par.sp if par is not None else None
        kw = dict(dit=dit, latents=latents, timestep=ts, y=y, controlnet=cn,
                  control_signal_video_latents=control_latents, sequence_parallel=sp)
        use_cfg = cfg_scale != 1.0
        if par is not None and par.layout.cfg_size == 2 and use_cfg:
            mine = self.model_fn(context=context_posi if par.layout.cfg_index == 0 else context_nega, **kw)
            both = torch.empty((2,) + tuple(mine.shape), dtype=mine.dtype, device=mine.device)
            par.dist.all_gather_into_tensor(both.view(-1), mine.contiguous().view(-1), group=par.cfg_group)
            posi, nega = both[0], both[1]
        else:
            posi = self.model_fn(context=context_posi, **kw)                           # :710
            nega = self.model_fn(context=context_nega, **kw) if use_cfg else None      # :715
        dsigma = self.scheduler.dsigma(timestep)                                       # flow_match.py:72-81
        return capi.cfg_euler(posi.contiguous(), None if nega is None else nega.contiguous(), latents.contiguous(),
                              cfg_scale, dsigma, out=out)                              # :716,721

    @torch.no_grad()
    def __call__(self, latents, context_posi, context_nega=None, y=None, control_latents=None,
                 num_inference_steps: int = 50, cfg_scale: float = 5.0, sigma_shift: float = 5.0,
                 switch_DiT_boundary: float = 0.875, denoising_strength: float = 1.0, callback=None):
        """Runs the whole schedule (defaults as wan_video_new.py:634-644) and returns the final latents."""
        self.scheduler.set_timesteps(num_inference_steps, denoising_strength=denoising_strength, shift=sigma_shift)
        if cfg_scale != 1.0 and context_nega is None:
            raise ValueError("cfg_scale != 1 needs a negative-prompt context")
        for i, t in enumerate(self.scheduler.timesteps):
            latents = self.step(latents, t, context_posi, context_nega, y, control_latents, cfg_scale,
                                switch_DiT_boundary)
            if callback is not None:
                callback(i, t, latents)
        return latents
