"""Pins the oracle (oracle/) against vectors produced by the unmodified reference (tests/golden/, made by
oracle/gen_golden.py), and -- when /root/reference is present -- against the reference itself, live."""
import json

import numpy as np
import pytest
import torch

from oracle import control_channels_oracle as CC
from oracle import wan_dit_oracle as O


def _cfg(d):
    d = dict(d)
    d["patch_size"] = tuple(d["patch_size"])
    return O.DiTConfig(**d)


@pytest.fixture(scope="module")
def dit_golden(golden_dir):
    return torch.load(golden_dir / "dit_forward.pt", weights_only=False)


@pytest.mark.parametrize("name", ["tiny_i2v", "tiny_t2v"])
def test_oracle_matches_reference_golden_fp32(dit_golden, name):
    g = dit_golden[name]
    cfg = _cfg(g["cfg"])
    sd = O.random_state_dict(cfg, seed=g["weight_seed"])
    inp = O.synthetic_inputs(cfg, *g["shape"], seed=g["input_seed"], ctx_len=g["ctx_len"], ctx_valid=g["ctx_valid"],
                             timestep=g["timestep"])
    with torch.no_grad():
        out = O.model_fn(sd, cfg, inp["latents"], inp["timestep"], inp["context"], y=inp.get("y"))
    assert out.shape == g["out_fp32"].shape
    # same torch ops in the same order: equal up to BLAS blocking differences between hosts
    assert O.rel_l2(out, g["out_fp32"]) < 1e-5


@pytest.mark.parametrize("name", ["tiny_i2v", "tiny_t2v"])
def test_oracle_matches_reference_golden_bf16(dit_golden, name):
    g = dit_golden[name]
    cfg = _cfg(g["cfg"])
    sd = O.random_state_dict(cfg, seed=g["weight_seed"], dtype=torch.bfloat16)
    inp = O.synthetic_inputs(cfg, *g["shape"], seed=g["input_seed"], ctx_len=g["ctx_len"], ctx_valid=g["ctx_valid"],
                             timestep=g["timestep"], dtype=torch.bfloat16)
    with torch.no_grad():
        out = O.model_fn(sd, cfg, inp["latents"], inp["timestep"], inp["context"], y=inp.get("y"))
    assert out.dtype == torch.bfloat16
    # bf16 accumulation order may differ between CPU generations; the golden's own distance to fp32 is ~1e-2
    assert O.rel_l2(out, g["out_bf16"]) < 2e-2
    assert O.rel_l2(out, g["out_fp32"]) < 5e-2


def test_oracle_controlnet_matches_reference_golden(dit_golden):
    g = dit_golden["a14b_slice_controlnet"]
    cfg = _cfg(g["cfg"])
    sd = O.random_state_dict(cfg, seed=g["weight_seed"])
    csd = O.random_controlnet_state_dict(cfg, 1, seed=g["controlnet_seed"])
    inp = O.synthetic_inputs(cfg, *g["shape"], seed=g["input_seed"], ctx_len=g["ctx_len"], ctx_valid=g["ctx_valid"],
                             timestep=g["timestep"])
    with torch.no_grad():
        base = O.model_fn(sd, cfg, inp["latents"], inp["timestep"], inp["context"], y=inp["y"])
        out = O.model_fn(sd, cfg, inp["latents"], inp["timestep"], inp["context"], y=inp["y"], controlnet_sd=csd,
                         control_signal_video_latents=inp["control_signal_video_latents"], controlnet_num_layers=1)
        czero = O.random_controlnet_state_dict(cfg, 1, seed=g["controlnet_seed"], zero_convs=True)
        noop = O.model_fn(sd, cfg, inp["latents"], inp["timestep"], inp["context"], y=inp["y"], controlnet_sd=czero,
                          control_signal_video_latents=inp["control_signal_video_latents"], controlnet_num_layers=1)
    assert O.rel_l2(base, g["out_base_fp32"]) < 1e-5
    assert O.rel_l2(out, g["out_fp32"]) < 1e-5
    assert O.rel_l2(out, base) > 1e-3            # the branch is visible with non-zero zero-convs
    assert torch.equal(noop, base)               # design invariant: untrained ControlNet == base model (bit-exact)


@pytest.mark.reference
def test_oracle_bit_identical_to_live_reference():
    from oracle import ref_shim
    ns = ref_shim.load()
    cfg = O.DiTConfig(dim=256, in_dim=36, ffn_dim=512, out_dim=16, text_dim=64, freq_dim=256, eps=1e-6, num_heads=2,
                      num_layers=2)
    sd = O.random_state_dict(cfg, seed=5)
    m = ns.WanModel(**ref_shim.cfg_kwargs(cfg)).eval()
    m.load_state_dict(sd, strict=True)
    inp = O.synthetic_inputs(cfg, 2, 4, 6, seed=6, ctx_len=16, ctx_valid=4, timestep=412.0)
    with torch.no_grad():
        ref = ns.model_fn_wan_video(dit=m, latents=inp["latents"], timestep=inp["timestep"], context=inp["context"],
                                    y=inp["y"])
        mine = O.model_fn(sd, cfg, inp["latents"], inp["timestep"], inp["context"], y=inp["y"])
    assert torch.equal(ref, mine)


def test_sinusoidal_embedding_rounds_timestep_like_the_pipeline():
    # F11: the pipeline casts the timestep to bf16 before the embedding (wan_video_new.py:707)
    t = torch.tensor([937.3], dtype=torch.bfloat16)
    e = O.sinusoidal_embedding_1d(256, t)
    assert e.dtype == torch.bfloat16 and e.shape == (1, 256)
    assert float(t) == 936.0
    assert abs(float(e[0, 0]) - float(torch.cos(torch.tensor(936.0, dtype=torch.float64)))) < 4e-3


def test_rope_table_split_and_unit_modulus():
    fr = O.rope_freqs(128, 3, 4, 5, "cpu")
    assert fr.shape == (60, 1, 64) and fr.dtype == torch.complex128
    assert torch.allclose(fr.abs(), torch.ones_like(fr.abs()))
    # token 0 is the identity rotation; the (f,h,w) axes own 22/21/21 complex pairs
    assert torch.allclose(fr[0], torch.ones_like(fr[0]))
    f1 = fr[4 * 5]          # f=1,h=0,w=0: only the first 22 pairs rotate
    assert not torch.allclose(f1[0, :22], torch.ones(22, dtype=torch.complex128))
    assert torch.allclose(f1[0, 22:], torch.ones(42, dtype=torch.complex128))


def test_control_channel_oracle_digests(golden_dir):
    g = json.loads((golden_dir / "control_channels.json").read_text())
    ranges = dict(min_force=30., max_force=400., min_indirect_force=30., max_indirect_force=400., min_mass=1.,
                  max_mass=4.)
    for name in ("_pendulum", "_golf", "_ballthendominos"):
        np.random.seed(0)
        cv = CC.control_video(**CC.row_to_args(g["rows"][name]), num_frames=81, height=480, width=832, **ranges)
        assert cv.shape == (81, 480, 832, 3) and cv.dtype == torch.bfloat16
        assert CC.digest(cv) == g["goal_force"][name]
        assert float(cv[..., 0].abs().max()) == 0.0      # goal-force rows: direct-force channel empty
