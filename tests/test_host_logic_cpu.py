"""Host-side logic of the product that runs without a GPU: scheduler table, control channels (bit-exact), mask
construction, parallel layout, RoPE table, weight re-layout."""
import json

import numpy as np
import pytest
import torch

from goal_force_b200 import control_channels as P
from goal_force_b200.pipeline import ParallelLayout, first_frame_mask, generate_noise, image_condition, latent_shape
from goal_force_b200.scheduler import FlowMatchScheduler
from goal_force_b200.wan_dit import rope_cos_sin
from oracle import control_channels_oracle as CC
from oracle import wan_dit_oracle as O


@pytest.mark.parametrize("steps", [40, 50, 4])
def test_scheduler_matches_reference_table(golden_dir, steps):
    g = json.loads((golden_dir / "scheduler.json").read_text())[f"{steps}_5.0"]
    s = FlowMatchScheduler(shift=5, sigma_min=0.0, extra_one_step=True)
    s.set_timesteps(steps, shift=5.0)
    assert s.sigmas.tolist() == g["sigmas"]              # bit-exact fp32 table
    assert s.timesteps.tolist() == g["timesteps"]
    assert int((s.timesteps >= 875).sum()) == g["n_high_noise"]
    # dsigma picks the same neighbours as the reference's step(); last step goes to sigma 0
    assert s.dsigma(s.timesteps[0]) == float(torch.tensor(g["sigmas"][1]) - torch.tensor(g["sigmas"][0]))
    assert s.dsigma(s.timesteps[-1]) == float(torch.zeros(()) - torch.tensor(g["sigmas"][-1]))


def test_expert_split_counts():
    # 40 steps -> 17 high-noise / 23 low-noise; 50 steps -> 21 / 29
    for steps, high in ((40, 17), (50, 21)):
        s = FlowMatchScheduler()
        s.set_timesteps(steps, shift=5.0)
        assert int((s.timesteps >= 0.875 * 1000).sum()) == high


def test_control_channels_bit_exact_all_examples(golden_dir):
    g = json.loads((golden_dir / "control_channels.json").read_text())
    assert len(g["goal_force"]) == 12
    for name, want in g["goal_force"].items():
        np.random.seed(0)
        cv = P.control_video_from_csv_row(g["rows"][name])
        assert CC.digest(cv) == want, name
    for name in ("_pendulum", "_toycar", "_cantaloupes", "_paw_tool2"):
        row = dict(g["rows"][name])
        row.update(projectile_force_magnitude=250.0, projectile_force_angle=37.0, projectile_mass=2.5,
                   target_indirect_force_magnitude=-1.0)
        np.random.seed(0)
        cv = P.control_video_from_csv_row(row)
        assert CC.digest(cv) == g["direct_force"][name], name
        assert float(cv[..., 1].abs().max()) == 0.0 and float(cv[..., 0].max()) > 0.9


def test_control_channels_both_forces_consumes_rng_like_reference():
    row = dict(projectile_force_magnitude=100.0, projectile_force_angle=10.0, projectile_coordx=100, projectile_coordy=50,
               projectile_mass=2.0, target_indirect_force_angle=200.0, target_indirect_force_magnitude=300.0,
               target_coordx=300, target_coordy=80, target_mass=3.0, width=208, height=120)
    kw = dict(num_frames=5, height=120, width=208, min_force=30., max_force=400., min_indirect_force=30.,
              max_indirect_force=400., min_mass=1., max_mass=4., p_mask_out_direct_force=0.3,
              p_mask_out_indirect_force=0.3, p_mask_out_masses=0.5)
    for seed in range(6):
        r1, r2 = np.random.RandomState(seed), np.random.RandomState(seed)
        a = CC.control_video(**CC.row_to_args(row), rng=r1, **kw)
        b = P.generate_control_video(P.ControlSignalSpec.from_csv_row(row), rng=r2, **kw)
        assert torch.equal(a, b)
        assert r1.uniform() == r2.uniform()              # same number of draws consumed


def test_first_frame_mask_and_condition():
    m = first_frame_mask(81, 6, 8)
    assert m.shape == (4, 21, 6, 8)
    assert float(m[:, 0].min()) == 1.0 and float(m[:, 1:].max()) == 0.0
    m2 = first_frame_mask(81, 6, 8, end_image=True)
    assert float(m2[3, -1].min()) == 1.0 and float(m2[:3, -1].max()) == 0.0
    y = image_condition(torch.randn(16, 21, 6, 8), 81)
    assert y.shape == (1, 20, 21, 6, 8)
    assert latent_shape(81, 480, 832) == (1, 16, 21, 60, 104)


def test_image_condition_matches_reference_unit_golden(golden_dir):
    """a21: y = cat(mask, vae_latents)[None] is bit-identical to what the reference's WanVideoUnit_ImageEmbedderVAE
    (src/goal_force/wan_video_new.py:887-917) builds, incl. the end-image variant and a 17-frame clip; vectors come
    from oracle/gen_golden.py::gen_mask, which runs the reference unit itself (VAE stubbed by seeded latents)."""
    g = torch.load(golden_dir / "image_condition.pt", weights_only=False)
    assert set(g) == {"small", "small_end", "odd_frames", "full"}
    for name, e in g.items():
        gen = torch.Generator("cpu").manual_seed(77)
        lat = torch.randn(16, (e["num_frames"] - 1) // 4 + 1, e["height"] // 8, e["width"] // 8, generator=gen)
        y = image_condition(lat.to(torch.bfloat16), e["num_frames"], end_image=e["end_image"])
        assert y.dtype == torch.bfloat16
        assert CC.digest(y) == e["digest"], name
        if "y" in e:
            assert torch.equal(y, e["y"]), name
            m = first_frame_mask(e["num_frames"], e["height"] // 8, e["width"] // 8, e["end_image"])
            assert torch.equal(m.to(torch.bfloat16), e["y"][0, :4]), name


def test_generate_noise_is_seed_reproducible_on_cpu():
    a = generate_noise((1, 16, 2, 4, 4), seed=5, device="cpu")
    g = torch.Generator("cpu").manual_seed(5)
    b = torch.randn((1, 16, 2, 4, 4), generator=g, dtype=torch.float32).to(torch.bfloat16)
    assert torch.equal(a, b)


def test_parallel_layout():
    lay = ParallelLayout(world_size=8, rank=5, cfg_size=2)
    assert (lay.sp_size, lay.cfg_index, lay.sp_index) == (4, 1, 1)
    assert lay.sp_ranks() == [4, 5, 6, 7] and lay.cfg_ranks() == [1, 5]
    with pytest.raises(ValueError):
        ParallelLayout(world_size=3, rank=0, cfg_size=2)


def test_rope_table_matches_oracle_freqs():
    f, h, w = 3, 4, 5
    cs = rope_cos_sin(128, f, h, w, "cpu")
    fr = O.rope_freqs(128, f, h, w, "cpu")[:, 0]
    assert cs.shape == (60, 64, 2)
    assert torch.equal(cs[..., 0], fr.real.float()) and torch.equal(cs[..., 1], fr.imag.float())
    sl = slice(20, 40)
    assert torch.equal(rope_cos_sin(128, f, h, w, "cpu", sl), cs[sl])


@pytest.mark.reference
def test_model_fn_keyword_surface_matches_reference():
    """Drop-in seam: every keyword the reference's model_fn_wan_video declares is accepted by ours with
    the same default, so `pipe.model_fn = goal_force_b200.wan_dit.model_fn_wan_video` needs no call-site change."""
    import inspect
    from oracle import ref_shim
    from goal_force_b200.wan_dit import model_fn_wan_video as ours
    ref = ref_shim.load().model_fn_wan_video
    rp, op = inspect.signature(ref).parameters, inspect.signature(ours).parameters
    assert any(p.kind is inspect.Parameter.VAR_KEYWORD for p in op.values())      # shared-input junk is swallowed
    for name, p in rp.items():
        if p.kind is inspect.Parameter.VAR_KEYWORD:
            continue
        if name in op and p.default is not inspect.Parameter.empty:
            assert op[name].default == p.default, name
    # the names the pipeline always passes (in_iteration_models + inputs) are explicit parameters on our side
    for name in ("dit", "controlnet", "latents", "timestep", "context", "y", "clip_feature"):
        assert name in op and name in rp, name
    wm_ref = inspect.signature(ref_shim.load().WanModel.forward).parameters
    from goal_force_b200.wan_dit import WanModelB200
    wm_ours = inspect.signature(WanModelB200.forward).parameters
    assert [n for n in wm_ref if n not in ("self", "kwargs")] == [n for n in wm_ours if n not in ("self", "kwargs")]


@pytest.mark.reference
def test_from_reference_modules_weight_layout():
    """WanModelB200.from_reference / ControlNetB200.from_reference read the live reference nn.Modules: config taken
    from the module attributes, q|k|v and cross k|v fused in the order the kernels expect, conv weights flattened with
    K index c*4 + kh*2 + kw, an untouched (zero-conv) ControlNet recognised as a no-op."""
    from oracle import ref_shim
    from goal_force_b200.wan_dit import ControlNetB200, WanModelB200
    ns = ref_shim.load()
    cfg = O.DiTConfig(dim=256, in_dim=36, ffn_dim=512, out_dim=16, text_dim=64, freq_dim=256, eps=1e-6, num_heads=2,
                      num_layers=2)
    torch.manual_seed(0)
    ref = ns.WanModel(**ref_shim.cfg_kwargs(cfg)).eval()
    ours = WanModelB200.from_reference(ref, device="cpu")
    c = ours.cfg
    assert (c.dim, c.in_dim, c.ffn_dim, c.num_heads, c.num_layers, c.out_dim) == (256, 36, 512, 2, 2, 16)
    sd = ref.state_dict()
    b = ours.blocks[1]
    want = torch.cat([sd[f"blocks.1.self_attn.{p}.weight"] for p in "qkv"], 0).bfloat16()
    assert torch.equal(b.wqkv, want)
    assert torch.equal(b.ckv_w, torch.cat([sd[f"blocks.1.cross_attn.{p}.weight"] for p in "kv"], 0).bfloat16())
    assert torch.equal(ours.patch_w, sd["patch_embedding.weight"].reshape(256, 36 * 4).bfloat16())
    assert torch.equal(ours.block_mod[0], sd["blocks.0.modulation"].reshape(-1).bfloat16())
    # the goal-force ControlNet hard-codes A14B widths; one layer keeps this CPU-sized
    cn_ref = ns.ControlNet(1, torch_dtype=torch.bfloat16)
    cn = ControlNetB200.from_reference(cn_ref, device="cpu")
    assert (cn.cfg.dim, cn.cfg.num_heads, cn.cfg.ffn_dim, cn.num_layers, cn.stride) == (5120, 40, 13824, 1, None)
    assert cn.is_noop                                          # zero_module(...) convs: branch is an exact no-op
    assert cn.patch_w.shape == (5120, 64) and cn.zero_w[0].shape == (5120, 5120)


@pytest.mark.reference
def test_standins_mirror_live_reference():
    """oracle/ref_standins.py (used on the GPU box, where /root/reference is absent) has the reference's state_dict keys,
    shapes and the attribute surface goal_force_b200 reads."""
    from oracle import ref_shim
    from oracle import ref_standins as S
    ns = ref_shim.load()
    cfg = O.DiTConfig(dim=256, in_dim=36, ffn_dim=512, out_dim=16, text_dim=64, freq_dim=256, eps=1e-6, num_heads=2,
                      num_layers=2)
    ref = ns.WanModel(**ref_shim.cfg_kwargs(cfg))
    mine = S.wan_standin_from_cfg(cfg, dtype=torch.float32)
    rs, ms = ref.state_dict(), mine.state_dict()
    assert list(rs) == list(ms)
    assert all(rs[k].shape == ms[k].shape for k in rs)
    for a in ("dim", "in_dim", "freq_dim", "has_image_input", "seperated_timestep", "require_vae_embedding",
              "require_clip_embedding", "fuse_vae_embedding_in_latents", "has_image_pos_emb", "has_ref_conv",
              "control_adapter"):
        assert getattr(ref, a) == getattr(mine, a), a
    assert tuple(ref.patch_size) == tuple(mine.patch_size)
    rb, mb = ref.blocks[0], mine.blocks[0]
    assert (rb.dim, rb.num_heads, rb.ffn_dim, rb.norm1.eps) == (mb.dim, mb.num_heads, mb.ffn_dim, mb.norm1.eps)
    assert ref.head.head.out_features == mine.head.head.out_features
    assert ref.text_embedding[0].in_features == mine.text_embedding[0].in_features
    cn_ref = ns.ControlNet(1, torch_dtype=torch.bfloat16)
    cn = S.ControlNetStandIn(1)
    crs, cms = cn_ref.state_dict(), cn.state_dict()
    assert list(crs) == list(cms)
    assert all(crs[k].shape == cms[k].shape and crs[k].dtype == cms[k].dtype for k in crs)
    assert (cn_ref.num_layers, cn_ref.stride) == (cn.num_layers, cn.stride)
    b0, m0 = cn_ref.controlnet_dit.blocks[0], cn.controlnet_dit.blocks[0]
    assert (b0.dim, b0.num_heads, b0.ffn_dim, b0.norm1.eps) == (m0.dim, m0.num_heads, m0.ffn_dim, m0.norm1.eps)


def test_conversion_cache_tracks_weight_changes(monkeypatch):
    """a reference module converted once must be re-converted when its weights change (load_state_dict /
    LoRA merge / load_controlnet_weights after a warm-up call), is_noop must follow, unsupported variants raise, and
    the cache must not keep the reference module alive."""
    import gc
    import weakref
    from goal_force_b200 import wan_dit as W
    from oracle import ref_standins as S
    monkeypatch.setattr(W, "_module_device", lambda m: torch.device("cpu"))
    monkeypatch.setattr(W.capi, "load", lambda: None)
    cfg = O.DiTConfig(dim=256, in_dim=36, ffn_dim=512, out_dim=16, text_dim=64, freq_dim=256, eps=1e-6, num_heads=2,
                      num_layers=1)
    cn = S.controlnet_standin_from_cfg(cfg, 1)
    # head_dim must be 128 for the kernels: dim 256 / 2 heads
    a = W._as_b200(cn, W.ControlNetB200)
    assert a.is_noop and W._as_b200(cn, W.ControlNetB200) is a            # cached while the weights are unchanged
    with torch.no_grad():
        cn.controlnet_zero_convs_after[0].weight.add_(0.01)               # what load_controlnet_weights does
    b = W._as_b200(cn, W.ControlNetB200)
    assert b is not a and not b.is_noop
    cn.load_state_dict({k: torch.zeros_like(v) for k, v in cn.state_dict().items()})
    assert W._as_b200(cn, W.ControlNetB200).is_noop
    ref = weakref.ref(cn)
    key = id(cn)
    del cn
    gc.collect()
    assert ref() is None and key not in W._CONVERTED                      # the cache holds the module only weakly
    dit = S.wan_standin_from_cfg(cfg)
    dit.seperated_timestep = True
    with pytest.raises(NotImplementedError, match="seperated_timestep"):
        W.WanModelB200.from_reference(dit, device="cpu")
