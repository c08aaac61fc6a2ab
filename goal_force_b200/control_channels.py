"""Goal Force / Direct Force control-channel construction (host side, bit-exact with the reference).

Mirrors ControlSignalDataset_Balls._generate_control_video / get_blob_for_mass / get_gaussian_blob
(src/goal_force/unified_dataset.py:775-940) and the validation-row handling of get_batch (:942-980):
    channel 0  moving Gaussian blob (radius 20) for the direct (projectile) force,
    channel 1  moving Gaussian blob for the goal (indirect, target) force,
    channel 2  static Gaussian blobs whose radius encodes the masses,
returned as an (F, H, W, 3) bf16 tensor in [0, 1].  In the reference this runs on the CPU inside the dataset
(`__getitem__`), once per CSV row, before the pipeline is called; it stays host code here.  All frames of a moving
blob are produced by one vectorised torch expression instead of a per-frame Python loop; the arithmetic per
element (int grid -> fp32 subtract, square, add, divide by 2 r^2, exp) is unchanged, so the bf16 result is
bit-identical (tests/test_host_logic_cpu.py and tests/test_jobs_cpu.py check SHA-256 digests produced by the reference's own class).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np
import torch

# scripts/inference/inference_goal_force.py:137-144: ranges the shipped checkpoint was trained with
INFERENCE_RANGES = dict(min_mass=1.0, max_mass=4.0, min_force=30.0, max_force=400.0,
                        min_indirect_force=30.0, max_indirect_force=400.0)


@dataclass
class ControlSignalSpec:
    """One CSV row in the units _generate_control_video expects (positions normalised to [0,1], y up)."""
    force: float
    angle: float
    x_pos: float
    y_pos: float
    target_indirect_force: float
    target_indirect_angle: float
    target_x_pos: float
    target_y_pos: float
    masses: dict = field(default_factory=lambda: {"projectile": -1, "target": -1, "distractors": []})
    coords: dict = field(default_factory=lambda: {"projectile": [0, 0], "target": [0, 0], "distractors": []})

    @classmethod
    def from_csv_row(cls, item) -> "ControlSignalSpec":
        """get_batch for an image (validation) row, unified_dataset.py:942-980."""
        return cls(
            force=item["projectile_force_magnitude"], angle=item["projectile_force_angle"],
            x_pos=item["projectile_coordx"] / item["width"], y_pos=item["projectile_coordy"] / item["height"],
            target_indirect_force=item["target_indirect_force_magnitude"],
            target_indirect_angle=item["target_indirect_force_angle"],
            target_x_pos=item["target_coordx"] / item["width"], target_y_pos=item["target_coordy"] / item["height"],
            masses={"projectile": item["projectile_mass"], "target": item["target_mass"], "distractors": []},
            coords={"projectile": [int(item["projectile_coordx"]), int(item["projectile_coordy"])],
                    "target": [int(item["target_coordx"]), int(item["target_coordy"])], "distractors": []})


def _grids(height: int, width: int):
    return torch.meshgrid(torch.arange(height), torch.arange(width), indexing="ij")


def _moving_blob(xs: float, ys: float, xe: float, ye: float, num_frames: int, height: int, width: int) -> torch.Tensor:
    """(F, H, W) fp32: frame i holds exp(-((X-x_i)^2 + (Y-y_i)^2) / (2*20^2)), (x_i, y_i) on the segment start->end.
    Centre coordinates are formed in Python doubles exactly as the reference does and rounded to fp32 at the
    tensor op, which is where `x_grid - x` rounds the Python scalar too."""
    cx = np.empty(num_frames, dtype=np.float64)
    cy = np.empty(num_frames, dtype=np.float64)
    for i in range(num_frames):
        t = i / (num_frames - 1)
        cx[i] = xs * (1 - t) + xe * t
        cy[i] = ys * (1 - t) + ye * t
    yy, xx = _grids(height, width)
    cxt = torch.from_numpy(cx).to(torch.float32).view(-1, 1, 1)
    cyt = torch.from_numpy(cy).to(torch.float32).view(-1, 1, 1)
    sq = (xx - cxt) ** 2 + (yy - cyt) ** 2
    return torch.exp(-sq / (2.0 * 20 ** 2))


def _static_blob(x, y, radius: float, height: int, width: int) -> torch.Tensor:
    yy, xx = _grids(height, width)
    sq = (xx - x) ** 2 + (yy - y) ** 2
    return torch.exp(-sq / (2.0 * radius ** 2))


def generate_control_video(spec: ControlSignalSpec, *, num_frames: int = 81, height: int = 480, width: int = 832,
                           min_force: float = 30.0, max_force: float = 400.0, min_indirect_force: float = 30.0,
                           max_indirect_force: float = 400.0, min_mass: float = 1.0, max_mass: float = 4.0,
                           p_mask_out_direct_force: float = 0.0, p_mask_out_indirect_force: float = 0.0,
                           p_mask_out_masses: float = 0.0, rng=np.random) -> torch.Tensor:
    """_generate_control_video (unified_dataset.py:775-889). `rng` is consumed exactly like np.random there."""
    force, tforce = spec.force, spec.target_indirect_force
    if force == -1:
        mask_direct, mask_indirect = True, False
    elif tforce == -1:
        mask_direct, mask_indirect = False, True
    else:
        mask_direct = mask_indirect = False
        u = rng.uniform(low=0.0, high=1.0)
        if u < p_mask_out_direct_force:
            mask_direct = True
        elif p_mask_out_direct_force <= u <= p_mask_out_direct_force + p_mask_out_indirect_force:
            mask_indirect = True
    d_max, d_min = width / 2, width / 8
    out = torch.zeros((num_frames, height, width, 3))

    def segment(mag, ang, xp, yp, lo, hi):
        xs, ys = xp * width, (1 - yp) * height
        total = d_min + (d_max - d_min) * ((mag - lo) / (hi - lo))
        return xs, ys, xs + total * math.cos(ang * torch.pi / 180.0), ys - total * math.sin(ang * torch.pi / 180.0)

    if not mask_direct:
        out[..., 0] = _moving_blob(*segment(force, spec.angle, spec.x_pos, spec.y_pos, min_force, max_force),
                                   num_frames, height, width)
    if not mask_indirect:
        out[..., 1] = _moving_blob(*segment(tforce, spec.target_indirect_angle, spec.target_x_pos, spec.target_y_pos,
                                            min_indirect_force, max_indirect_force), num_frames, height, width)
    mask_masses = rng.uniform(low=0.0, high=1.0) < p_mask_out_masses        # always drawn (:851)
    if not mask_masses:
        def radius(m):
            t = (m - min_mass) / (max_mass - min_mass)
            return (1 - t) * 5 + t * 40

        ch2 = torch.zeros((height, width))
        any_mass = False
        items = [(spec.masses["projectile"], spec.coords["projectile"]), (spec.masses["target"], spec.coords["target"])]
        items += [(m, c) for m, c in zip(spec.masses["distractors"], spec.coords["distractors"])]
        for k, (m, (cx, cy)) in enumerate(items):
            if (k < 2 and m > -1) or (k >= 2 and m != -1):
                ch2 = ch2 + _static_blob(cx, height - cy, radius(m), height, width)
                any_mass = True
        if any_mass:
            out[..., 2] = ch2
        out = torch.clamp(out, min=0.0, max=1.0)                              # only in this branch (:887)
    return out.to(torch.bfloat16)


def control_video_from_csv_row(item, *, num_frames: int = 81, height: int = 480, width: int = 832,
                               rng=np.random, **ranges) -> torch.Tensor:
    """CSV row (dict / pandas Series) -> control video with the inference-time ranges of the shipped script."""
    r = dict(INFERENCE_RANGES)
    r.update(ranges)
    return generate_control_video(ControlSignalSpec.from_csv_row(item), num_frames=num_frames, height=height,
                                  width=width, rng=rng, **r)
