"""Restatement of the Goal Force control-channel synthesis -- TEST INFRASTRUCTURE ONLY (see oracle/wan_dit_oracle.py).

Follows src/goal_force/unified_dataset.py:775-940 (`_generate_control_video`, `get_blob_for_mass`,
`get_gaussian_blob`; Balls and Dominos variants are identical) and the CSV row handling of `get_batch`
(:942-1027) for validation (image) rows. Written with explicit torch CPU ops so that the result is bit-identical to
the reference on the same machine; pinned by the SHA-256 digests in tests/golden/control_channels.json, which
oracle/gen_golden.py produces by running the reference's own dataset class.
"""
from __future__ import annotations

import math

import numpy as np
import torch


def gaussian_blob(x: float, y: float, radius: float, height: int, width: int) -> torch.Tensor:
    """get_gaussian_blob (unified_dataset.py:903-940) for one channel: exp(-((X-x)^2+(Y-y)^2) / (2 r^2)), fp32,
    on integer pixel grids."""
    yy, xx = torch.meshgrid(torch.arange(height), torch.arange(width), indexing="ij")
    sq = (xx - x) ** 2 + (yy - y) ** 2
    return 1.0 * torch.exp(-sq / (2.0 * radius ** 2))


def mass_blob(xpos, ypos, mass, min_mass, max_mass, height, width) -> torch.Tensor:
    """get_blob_for_mass (unified_dataset.py:891-901): radius 5 + 35 * (mass - min) / (max - min)."""
    t = (mass - min_mass) / (max_mass - min_mass)
    radius = (1 - t) * 5 + t * 40
    return gaussian_blob(xpos, ypos, radius, height, width)


def control_video(force, angle, x_pos, y_pos, target_force, target_angle, target_x_pos, target_y_pos, *,
                  num_frames, height, width, masses, coords, min_force, max_force, min_indirect_force,
                  max_indirect_force, min_mass, max_mass, p_mask_out_direct_force=0.0, p_mask_out_indirect_force=0.0,
                  p_mask_out_masses=0.0, rng=np.random) -> torch.Tensor:
    """_generate_control_video (unified_dataset.py:775-889) -> (F, H, W, 3) bf16.
    ch0: moving direct-force blob, ch1: moving goal(indirect)-force blob, ch2: static mass blobs.
    Draws from `rng` exactly where the reference draws from np.random (one uniform only when both forces are given,
    one uniform always for the mass mask)."""
    sig = torch.zeros((num_frames, 3, height, width))
    if force == -1:                                     # :784-801
        mask_direct, mask_indirect = True, False
    elif target_force == -1:
        mask_direct, mask_indirect = False, True
    else:
        mask_direct = mask_indirect = False
        u = rng.uniform(low=0.0, high=1.0)
        if u < p_mask_out_direct_force:
            mask_direct = True
        elif p_mask_out_direct_force <= u <= p_mask_out_direct_force + p_mask_out_indirect_force:
            mask_indirect = True
    disp_max, disp_min = width / 2, width / 8           # :803-804

    def moving(channel, f_mag, f_ang, xp, yp, fmin, fmax):
        xs, ys = xp * width, (1 - yp) * height
        pct = (f_mag - fmin) / (fmax - fmin)
        total = disp_min + (disp_max - disp_min) * pct
        xe = xs + total * math.cos(f_ang * torch.pi / 180.0)
        ye = ys - total * math.sin(f_ang * torch.pi / 180.0)
        for fr in range(num_frames):
            t = fr / (num_frames - 1)
            sig[:, channel][fr] += gaussian_blob(xs * (1 - t) + xe * t, ys * (1 - t) + ye * t, 20, height, width)

    if not mask_direct:                                  # :806-822
        moving(0, force, angle, x_pos, y_pos, min_force, max_force)
    if not mask_indirect:                                # :825-839
        moving(1, target_force, target_angle, target_x_pos, target_y_pos, min_indirect_force, max_indirect_force)
    sig = sig.permute(0, 2, 3, 1).contiguous()           # f c h w -> f h w c  (:842)
    sig[:, :, :, 2] = 0                                  # :848
    mask_masses = rng.uniform(low=0.0, high=1.0) < p_mask_out_masses      # :851 (always drawn)
    if not mask_masses:
        if masses["projectile"] > -1:
            sig[:, :, :, 2] += mass_blob(coords["projectile"][0], height - coords["projectile"][1],
                                         masses["projectile"], min_mass, max_mass, height, width)
        if masses["target"] > -1:
            sig[:, :, :, 2] += mass_blob(coords["target"][0], height - coords["target"][1], masses["target"],
                                         min_mass, max_mass, height, width)
        for m, (cx, cy) in zip(masses["distractors"], coords["distractors"]):
            if m == -1:
                continue
            sig[:, :, :, 2] += mass_blob(cx, height - cy, m, min_mass, max_mass, height, width)
        sig = torch.clamp(sig, min=0.0, max=1.0)         # :887 (only inside this branch)
    return sig.to(torch.bfloat16)


def row_to_args(item: dict) -> dict:
    """get_batch for a validation (image) CSV row (unified_dataset.py:942-980): normalised positions, int coords."""
    return dict(
        force=item["projectile_force_magnitude"], angle=item["projectile_force_angle"],
        x_pos=item["projectile_coordx"] / item["width"], y_pos=item["projectile_coordy"] / item["height"],
        target_force=item["target_indirect_force_magnitude"], target_angle=item["target_indirect_force_angle"],
        target_x_pos=item["target_coordx"] / item["width"], target_y_pos=item["target_coordy"] / item["height"],
        masses={"projectile": item["projectile_mass"], "target": item["target_mass"], "distractors": []},
        coords={"projectile": [int(item["projectile_coordx"]), int(item["projectile_coordy"])],
                "target": [int(item["target_coordx"]), int(item["target_coordy"])], "distractors": []},
    )


def digest(t: torch.Tensor) -> str:
    import hashlib
    return hashlib.sha256(t.contiguous().view(torch.int16).numpy().tobytes()).hexdigest()[:16]
