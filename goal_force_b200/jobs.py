"""Batched job driver for Goal Force / Direct Force inference (BASELINE.json configs[3]).

Replaces the per-CSV Python loop of scripts/inference/inference_goal_force.py:119-215 on the denoising side:

  * CSV rows -> jobs.  `shard_contiguous` is the reference's contiguous split of the example list over devices
    (scripts/inference/utils.py:25-57), here applied to replica groups instead of single GPUs.
  * The control video of a row is synthesised on the HOST with goal_force_b200.control_channels (bit-exact with the
    reference's ControlSignalDataset, including the np.random draws), in a background thread, while the GPUs denoise
    the previous row: 0.4 s of CPU work hides completely behind >= 16 s of denoising, so the control-channel
    construction stays bit-exact AND off the critical path.  (A GPU kernel cannot be bit-exact here: the reference's
    torch.exp on CPU is MKL's vsExp, whose rounding differs from a correctly rounded exp on ~1 % of the inputs;
    see DESIGN.md.)
  * Everything outside the hot path is a callable supplied by the caller, exactly where the reference calls its
    PyTorch VAE / T5 (north star: they may remain reference PyTorch): `encode_control(control_video) -> latents`,
    `conditioning(row) -> dict(context_posi, context_nega, y)`.
  * The (replicas x cfg x sequence-parallel) layout is a ParallelLayout; 8 GPUs run either one job at a time as
    cfg 2 x ulysses 4, or four jobs at a time as 4 replicas x cfg 2, etc.  The reference never batches rows
    (its collate returns batch[0]): four CSV rows are four independent B = 1 problems.
"""
from __future__ import annotations

import csv
import threading
from dataclasses import dataclass, field
from typing import Callable

import numpy as np
import torch

from . import control_channels as cc
from .pipeline import GoalForceDenoiser, ParallelContext, generate_noise, latent_shape

_NUMERIC = ("projectile_force_magnitude", "projectile_force_angle", "projectile_coordx", "projectile_coordy",
            "projectile_mass", "target_indirect_force_magnitude", "target_indirect_force_angle", "target_coordx",
            "target_coordy", "target_mass", "width", "height")


def shard_contiguous(items, world_size: int, device_id: int):
    """split_list_across_devices_contiguous (scripts/inference/utils.py:25-57): the first n % world_size shards get
    one extra item; shards are contiguous and in order."""
    n = len(items)
    base, rem = divmod(n, world_size)
    start = device_id * (base + 1) if device_id < rem else rem * (base + 1) + (device_id - rem) * base
    return items[start:start + base + (1 if device_id < rem else 0)]


def read_rows(csv_path) -> list:
    """CSV rows as dicts with the numeric columns converted (what pandas hands to the reference's get_batch)."""
    rows = []
    with open(csv_path, newline="") as f:
        for r in csv.DictReader(f):
            rows.append({k: (float(v) if k in _NUMERIC and v not in ("", None) else v) for k, v in r.items()})
    return rows


def direct_force_row(row: dict, magnitude: float, angle: float, mass: float) -> dict:
    """Direct Force variant of a CSV row (BASELINE.json configs[3]): the projectile gets a force and a mass, the goal
    force is switched off -- channel 0 carries the moving blob, channel 1 stays empty, channel 2 the masses."""
    r = dict(row)
    r.update(projectile_force_magnitude=float(magnitude), projectile_force_angle=float(angle),
             projectile_mass=float(mass), target_indirect_force_magnitude=-1.0)
    return r


@dataclass
class Job:
    row: dict
    seed: int = 0
    index: int = 0
    control_video: torch.Tensor | None = None          # (F, H, W, 3) bf16, host
    result: torch.Tensor | None = None                 # final latents (1, 16, T, H/8, W/8)
    video: torch.Tensor | None = None                  # decoded frames (1, 3, F, H, W) in [-1, 1] when a decoder is given
    info: dict = field(default_factory=dict)


class ControlVideoPrefetcher:
    """Builds the control video of job k+1 on a host thread while job k is on the GPUs.  np.random is seeded per job
    exactly as the shipped script seeds it per process (`np.random.seed(seed)` semantics are the caller's choice:
    pass `rng_seed`), and every job gets its own RandomState so the thread never touches the global generator."""

    def __init__(self, jobs, num_frames: int, height: int, width: int, rng_seed: int | None = 0):
        self.jobs, self.dims, self.rng_seed = jobs, (num_frames, height, width), rng_seed
        self._thread = None
        self._next = 0

    def _build(self, job: Job) -> None:
        f, h, w = self.dims
        rng = np.random.RandomState(self.rng_seed) if self.rng_seed is not None else np.random
        job.control_video = cc.control_video_from_csv_row(job.row, num_frames=f, height=h, width=w, rng=rng)

    def start(self, k: int) -> None:
        if k < len(self.jobs) and self.jobs[k].control_video is None:
            self._thread = threading.Thread(target=self._build, args=(self.jobs[k],), daemon=True)
            self._thread.start()

    def get(self, k: int) -> torch.Tensor:
        if self._thread is not None:
            self._thread.join()
            self._thread = None
        if self.jobs[k].control_video is None:
            self._build(self.jobs[k])
        return self.jobs[k].control_video


class BatchDriver:
    """Runs a list of CSV rows through the denoiser on this rank's replica group.

    denoiser       GoalForceDenoiser (already bound to this rank's ParallelContext)
    encode_control control video (F, H, W, 3) bf16 on the host -> control latents (1, 16, T, H/8, W/8) on the device;
                   in the reference this is preprocess + VAE encode (wan_video_new.py:791-805)
    conditioning   row -> dict(context_posi=, context_nega=, y=) on the device (T5 prompt embeddings and the
                   mask + VAE image latents, wan_video_new.py:808-820, 887-917)
    """

    def __init__(self, denoiser: GoalForceDenoiser, encode_control: Callable, conditioning: Callable,
                 parallel: ParallelContext | None = None, num_frames: int = 81, height: int = 480, width: int = 832,
                 num_inference_steps: int = 50, cfg_scale: float = 5.0, sigma_shift: float = 5.0, device="cuda",
                 decode: Callable | None = None):
        """decode: optional final latents (1, 16, T, H/8, W/8) -> video (1, 3, F, H, W) in [-1, 1]; in the reference
        this is `vae.decode` at the end of the pipeline call (wan_video_new.py:731-734), see `vae_decoder`."""
        self.decode = decode
        self.denoiser, self.encode_control, self.conditioning = denoiser, encode_control, conditioning
        self.parallel = parallel
        self.dims = (num_frames, height, width)
        self.steps, self.cfg_scale, self.sigma_shift = num_inference_steps, cfg_scale, sigma_shift
        self.device = device

    def my_jobs(self, rows, seed: int = 0) -> list:
        """The rows this rank's replica group works on (contiguous shards, like the reference's device split)."""
        lay = self.parallel.layout if self.parallel is not None else None
        replicas, replica = (lay.replicas, lay.replica) if lay is not None else (1, 0)
        idx = shard_contiguous(list(range(len(rows))), replicas, replica)
        return [Job(row=rows[i], seed=seed, index=i) for i in idx]

    @torch.no_grad()
    def run(self, rows, seed: int = 0, rng_seed: int | None = 0, on_done: Callable | None = None) -> list:
        jobs = self.my_jobs(rows, seed)
        f, h, w = self.dims
        pre = ControlVideoPrefetcher(jobs, f, h, w, rng_seed)
        pre.start(0)
        for k, job in enumerate(jobs):
            video = pre.get(k)
            pre.start(k + 1)                                   # next row's synthesis overlaps this row's denoising
            control = self.encode_control(video)
            cond = self.conditioning(job.row)
            noise = generate_noise(latent_shape(f, h, w), seed=job.seed, device=self.device)    # :751-763, CPU generator
            job.result = self.denoiser(noise, cond["context_posi"], cond.get("context_nega"), y=cond.get("y"),
                                       control_latents=control, num_inference_steps=self.steps,
                                       cfg_scale=self.cfg_scale, sigma_shift=self.sigma_shift)
            if self.decode is not None:
                job.video = self.decode(job.result)
            job.info = {"row": job.index, "control_digest_channels": [bool(video[..., c].any()) for c in range(3)]}
            if on_done is not None:
                on_done(job)
        return jobs


def vae_control_encoder(vae, tiled: bool = True, tile_size=(30, 52), tile_stride=(15, 26)):
    """`encode_control` on the B200 VAE (goal_force_b200.wan_vae.WanVideoVAEB200): what
    WanVideoUnit_ControlVideoEmbedder.process does (src/goal_force/wan_video_new.py:798-805) -- the (F, H, W, 3) control
    video goes in as it is, rearranged to (1, 3, F, H, W), and comes back as (1, 16, T, H/8, W/8) bf16 latents; the
    tiling defaults are the pipeline's (:648-650)."""

    def encode(video: torch.Tensor) -> torch.Tensor:
        v = video.to(device=vae.device, dtype=torch.bfloat16).permute(3, 0, 1, 2).unsqueeze(0)
        return vae.encode(v, vae.device, tiled=tiled, tile_size=tile_size, tile_stride=tile_stride)

    return encode


def vae_decoder(vae, tiled: bool = True, tile_size=(30, 52), tile_stride=(15, 26), group=None):
    """`decode` on the B200 VAE: the pipeline's last stage (src/goal_force/wan_video_new.py:731-734) with its tiling
    defaults.  `group`: the replica group that denoised the video -- its ranks then share the tiles (bit-identical
    result on every rank)."""

    def decode(latents: torch.Tensor) -> torch.Tensor:
        return vae.decode(latents, vae.device, tiled=tiled, tile_size=tile_size, tile_stride=tile_stride, group=group)

    return decode


def video_to_uint8(video: torch.Tensor) -> torch.Tensor:
    """vae_output_to_video (diffsynth/utils/__init__.py:76-91) without the PIL step: (1, 3, F, H, W) in [-1, 1] ->
    (F, H, W, 3) uint8, ((x + 1) * 127.5).clip(0, 255) truncated like the reference's `.to(torch.uint8)`."""
    frames = video[0].permute(1, 2, 3, 0).float()
    return ((frames + 1.0) * (255.0 / 2.0)).clip(0, 255).to(torch.uint8)


def vae_image_condition(vae, image: torch.Tensor, num_frames: int, tiled: bool = True, tile_size=(30, 52),
                        tile_stride=(15, 26)) -> torch.Tensor:
    """`y` of WanVideoUnit_ImageEmbedderVAE.process (src/goal_force/wan_video_new.py:894-916) on the B200 VAE:
    image (3, H, W) in [-1, 1] -> first frame of an otherwise zero clip -> VAE encode -> cat(mask, latents)[None]."""
    from .pipeline import image_condition
    c, h, w = image.shape
    clip = torch.zeros((c, num_frames, h, w), dtype=torch.bfloat16, device=vae.device)
    clip[:, 0] = image.to(device=vae.device, dtype=torch.bfloat16)
    lat = vae.encode([clip], vae.device, tiled=tiled, tile_size=tile_size, tile_stride=tile_stride)[0]
    return image_condition(lat, num_frames)


def synthetic_control_encoder(device="cuda"):
    """Stand-in for `VAE.encode(control video)` where no VAE weights exist (benchmarks, tests): a fixed linear map
    with the VAE's shape contract, (F, H, W, 3) in [0, 1] -> (1, 16, (F-1)/4+1, H/8, W/8) bf16: 8x8 spatial / 4-frame
    temporal average pooling of the three channels (first frame on its own, as the causal VAE does), scaled to
    [-1, 1] and mixed into 16 channels.  Deterministic, so a control video maps to the same latents on every rank."""
    mix = torch.linspace(-1.0, 1.0, 16 * 3).reshape(16, 3)

    def encode(video: torch.Tensor) -> torch.Tensor:
        v = video.to(device=device, dtype=torch.float32).permute(3, 0, 1, 2) * 2 - 1        # (3, F, H, W)
        sp = torch.nn.functional.avg_pool2d(v, 8)                                            # (3, F, H/8, W/8)
        first, rest = sp[:, :1], sp[:, 1:]
        rest = rest.reshape(3, (rest.shape[1]) // 4, 4, *rest.shape[2:]).mean(2)
        lat = torch.cat([first, rest], 1)                                                    # (3, T, h, w)
        out = torch.einsum("oc,cthw->othw", mix.to(device), lat) * 2.0
        return out.unsqueeze(0).to(torch.bfloat16).contiguous()

    return encode
