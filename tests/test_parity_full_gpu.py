"""Parity at the configurations bench.py measures ("make parity green on what you benchmark").

  * configs[1] in full: A14B, 40 trunk blocks + 10-block goal-force ControlNet, 81x480x832 = 32,760 tokens, ours vs the
    oracle on the same GPU in fp32 and in bf16.  Weights are generated lazily on the device (LazyRandomStateDict: the
    reference's key names and shapes, one tensor alive at a time), the oracle sees the very same bf16 values (upcast
    for its fp32 run), so 14 B parameters never have to exist three times.
  * the BASELINE.json sampling criterion (fixed-seed 40-step two-expert CFG run, final-latent cosine >= 0.999) at A14B
    width and the full 32,760-token length; depth is reduced to keep the suite short -- the full-depth run is
    tools/parity_full.py (log under profiles/).
  * strided ControlNet inject (src/goal_force/wan_video_new.py:1559-1563) against the oracle's controlnet_stride.
  * the INTEGRATION.md drop-in recipe with nn.Modules (reference attribute tree) handed to model_fn_wan_video.

Tolerance:  relL2(ours, ref_fp32) <= max(1e-2, 1.0 * relL2(ref_bf16, ref_fp32)).
"""
import pytest
import torch

from oracle import wan_dit_oracle as O

pytestmark = pytest.mark.gpu

FRAMES_LAT, H_LAT, W_LAT = 21, 60, 104          # 81 x 480 x 832 video -> 21*30*52 = 32,760 tokens


class CastView:
    """state-dict view: base[key] cast to `dtype` on access (nothing is stored)."""

    def __init__(self, base, dtype):
        self.base, self.dtype = base, dtype

    def __getitem__(self, key):
        return self.base[key].to(self.dtype)


@pytest.fixture(scope="module", autouse=True)
def _no_tf32(lib):
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    yield
    torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old
    torch.cuda.empty_cache()


def _prod_cfg(cfg):
    from goal_force_b200.wan_dit import DiTConfig
    return DiTConfig(**cfg.__dict__)


def _bound(e_ours, e_ref):
    return e_ours <= max(1e-2, 1.0 * e_ref)


def _free():
    import gc
    from goal_force_b200 import wan_dit
    wan_dit._WORKSPACES.clear()
    gc.collect()
    torch.cuda.empty_cache()


def test_config2_full_depth_full_length_vs_oracle():
    """The forward bench.py times -- 40 + 10 blocks, L = 32,760, ControlNet with non-zero zero-convs -- against the
    oracle in fp32 and bf16 on this device."""
    from goal_force_b200.synthetic import LazyRandomStateDict, synthetic_inputs
    from goal_force_b200.wan_dit import ControlNetB200, WanModelB200, model_fn_wan_video
    cfg, n_cn = O.WAN22_I2V_A14B, 10
    pc = _prod_cfg(cfg)
    lazy = LazyRandomStateDict(pc, seed=0, device="cuda")
    lazy_cn = LazyRandomStateDict(pc, seed=1, device="cuda", controlnet_layers=n_cn)
    inp = synthetic_inputs(pc, FRAMES_LAT, H_LAT, W_LAT, seed=1, device="cuda", timestep=937.0)
    dit = WanModelB200(pc, lazy)
    cn = ControlNetB200(pc, lazy_cn, n_cn)
    assert not cn.is_noop
    out = model_fn_wan_video(dit=dit, controlnet=cn, latents=inp["latents"], timestep=inp["timestep"],
                             context=inp["context"], y=inp["y"],
                             control_signal_video_latents=inp["control_signal_video_latents"])
    base = model_fn_wan_video(dit=dit, latents=inp["latents"], timestep=inp["timestep"], context=inp["context"],
                              y=inp["y"])
    torch.cuda.synchronize()
    assert out.shape == (1, 16, FRAMES_LAT, H_LAT, W_LAT) and not torch.isnan(out).any()
    assert O.rel_l2(out, base) > 1e-3                       # the ControlNet branch is visible at full depth
    out = out.cpu()
    del dit, cn, base
    _free()
    refs = []
    for dt in (torch.bfloat16, torch.float32):
        i = {k: v.to(dt) for k, v in inp.items()}
        with torch.no_grad():
            r = O.model_fn(CastView(lazy, dt), cfg, i["latents"], i["timestep"], i["context"], y=i["y"],
                           controlnet_sd=CastView(lazy_cn, dt),
                           control_signal_video_latents=i["control_signal_video_latents"], controlnet_num_layers=n_cn)
        refs.append(r.cpu())
        del r
        _free()
    refbf, ref32 = refs
    e_ours, e_ref, e_bf = O.rel_l2(out, ref32), O.rel_l2(refbf, ref32), O.rel_l2(out, refbf)
    print(f"config 2 full depth (40+10 blocks, 32760 tokens): relL2 ours-vs-fp32 {e_ours:.3e}  "
          f"ref_bf16-vs-fp32 {e_ref:.3e}  ours-vs-ref_bf16 {e_bf:.3e}  cosine-vs-fp32 {O.cosine(out, ref32):.6f}")
    assert _bound(e_ours, e_ref), (e_ours, e_ref)


def _sampler_case(layers, n_cn, steps=40):
    """ours vs the same loop written with the oracle forward in bf16 (wan_video_new.py:697-721); returns cosine."""
    from goal_force_b200.pipeline import GoalForceDenoiser, generate_noise
    from goal_force_b200.scheduler import FlowMatchScheduler
    from goal_force_b200.synthetic import LazyRandomStateDict, synthetic_inputs
    from goal_force_b200.wan_dit import ControlNetB200, WanModelB200
    cfg = O.DiTConfig(**{**O.WAN22_I2V_A14B.__dict__, "num_layers": layers})
    pc = _prod_cfg(cfg)
    lazies = [LazyRandomStateDict(pc, seed=s, device="cuda") for s in (20, 21)]
    lazy_cn = LazyRandomStateDict(pc, seed=22, device="cuda", controlnet_layers=n_cn)
    lazy_cn0 = LazyRandomStateDict(pc, seed=23, device="cuda", controlnet_layers=n_cn, zero_convs=True)
    inp = synthetic_inputs(pc, FRAMES_LAT, H_LAT, W_LAT, seed=24, device="cuda")
    g = torch.Generator("cpu").manual_seed(25)
    ctx_n = torch.randn(1, 512, cfg.text_dim, generator=g).to("cuda", torch.bfloat16)
    noise = generate_noise(tuple(inp["latents"].shape), seed=5)
    den = GoalForceDenoiser(WanModelB200(pc, lazies[0]), WanModelB200(pc, lazies[1]),
                            ControlNetB200(pc, lazy_cn, n_cn), ControlNetB200(pc, lazy_cn0, n_cn))
    assert den.controlnet2.is_noop and not den.controlnet.is_noop          # as the shipped inference script
    got = den(noise, inp["context"], ctx_n, y=inp["y"], control_latents=inp["control_signal_video_latents"],
              num_inference_steps=steps, cfg_scale=5.0, sigma_shift=5.0).cpu()
    del den
    _free()
    sch = FlowMatchScheduler()
    sch.set_timesteps(steps, shift=5.0)
    sds = [CastView(l, torch.bfloat16) for l in lazies]
    cnb = CastView(lazy_cn, torch.bfloat16)
    lat, used = noise, []
    with torch.no_grad():
        for t in sch.timesteps:
            e = 1 if float(t) < 875 else 0
            used.append(e)
            ts = t.unsqueeze(0).to("cuda", torch.bfloat16)
            kw = dict(y=inp["y"])
            if e == 0:      # controlnet2's zero-convs are all zero: the reference result equals the plain forward
                kw.update(controlnet_sd=cnb, control_signal_video_latents=inp["control_signal_video_latents"],
                          controlnet_num_layers=n_cn)
            p = O.model_fn(sds[e], cfg, lat, ts, inp["context"], **kw)
            n = O.model_fn(sds[e], cfg, lat, ts, ctx_n, **kw)
            pred = n + 5.0 * (p - n)
            s0, s1 = sch.sigma_pair(t)
            lat = lat + pred * (s1 - s0)
    return O.cosine(got, lat), O.rel_l2(got, lat), used


def test_sampler_40_steps_a14b_width_full_length_cosine():
    """BASELINE.json criterion at A14B width (d 5120, 40 heads, ffn 13824) and the full 32,760-token length: 40 steps,
    two experts (17 high-noise + 23 low-noise), CFG 5.0, shift 5.0, goal-force ControlNet on the high-noise expert.
    Depth is 4 trunk + 2 ControlNet blocks per expert here (the full 40 + 10 run is tools/parity_full.py)."""
    cos, rel, used = _sampler_case(layers=4, n_cn=2)
    assert used.count(0) == 17 and used.count(1) == 23
    print(f"40-step sampler, A14B width, 32760 tokens, 4+2 blocks: cosine {cos:.6f} relL2 {rel:.3e}")
    assert cos >= 0.999


def test_strided_controlnet_vs_oracle():
    """ControlNet(stride=2): the raw block states are added after trunk blocks 0 and 2 (no zero-conv),
    src/goal_force/wan_video_new.py:1559-1563; A14B widths, 4 trunk + 2 ControlNet blocks, 2080 tokens."""
    from goal_force_b200.wan_dit import ControlNetB200, WanModelB200, model_fn_wan_video
    cfg = O.DiTConfig(**{**O.WAN22_I2V_A14B.__dict__, "num_layers": 4})
    sd = O.random_state_dict(cfg, seed=40)
    csd = O.random_controlnet_state_dict(cfg, 2, seed=41)
    inp = O.synthetic_inputs(cfg, 4, 40, 52, seed=42, timestep=960.0)
    outs = []
    for dt in (torch.float32, torch.bfloat16):
        s = {k: v.to("cuda", dt) for k, v in sd.items()}
        c = {k: v.to("cuda", dt) for k, v in csd.items()}
        i = {k: v.to("cuda", dt) for k, v in inp.items()}
        with torch.no_grad():
            outs.append(O.model_fn(s, cfg, i["latents"], i["timestep"], i["context"], y=i["y"], controlnet_sd=c,
                                   control_signal_video_latents=i["control_signal_video_latents"],
                                   controlnet_num_layers=2, controlnet_stride=2))
            if dt == torch.float32:
                plain32 = O.model_fn(s, cfg, i["latents"], i["timestep"], i["context"], y=i["y"], controlnet_sd=c,
                                     control_signal_video_latents=i["control_signal_video_latents"],
                                     controlnet_num_layers=2)
        del s, c
    ref32, refbf = outs
    pc = _prod_cfg(cfg)
    bf = {k: v.to("cuda", torch.bfloat16) for k, v in inp.items()}
    cn = ControlNetB200(pc, csd, 2, stride=2)
    assert not cn.is_noop                                    # a strided ControlNet never uses its zero-convs
    out = model_fn_wan_video(dit=WanModelB200(pc, sd), controlnet=cn, latents=bf["latents"], timestep=bf["timestep"],
                             context=bf["context"], y=bf["y"],
                             control_signal_video_latents=bf["control_signal_video_latents"])
    e_ours, e_ref = O.rel_l2(out, ref32), O.rel_l2(refbf, ref32)
    print(f"strided ControlNet: relL2 ours-vs-fp32 {e_ours:.3e} ref_bf16-vs-fp32 {e_ref:.3e}")
    assert O.rel_l2(ref32, plain32) > 1e-2                   # the strided path really differs from the zero-conv path
    assert _bound(e_ours, e_ref), (e_ours, e_ref)


def test_long_sequence_75600_tokens_vs_oracle():
    """BASELINE configs[4] length: 81 x 720 x 1280 -> latent 21 x 90 x 160 -> 75,600 tokens (590 two-tile work items per
    head: the attention tail splitting and the 2.3x longer kv loop are exercised), A14B widths, 2 trunk + 1 ControlNet
    block, against the oracle on this device."""
    from goal_force_b200.wan_dit import ControlNetB200, WanModelB200, model_fn_wan_video
    cfg = O.DiTConfig(**{**O.WAN22_I2V_A14B.__dict__, "num_layers": 2})
    sd = O.random_state_dict(cfg, seed=70)
    csd = O.random_controlnet_state_dict(cfg, 1, seed=71)
    inp = O.synthetic_inputs(cfg, 21, 90, 160, seed=72, timestep=937.0)
    outs = []
    for dt in (torch.float32, torch.bfloat16):
        s_ = {k: v.to("cuda", dt) for k, v in sd.items()}
        c_ = {k: v.to("cuda", dt) for k, v in csd.items()}
        i = {k: v.to("cuda", dt) for k, v in inp.items()}
        with torch.no_grad():
            outs.append(O.model_fn(s_, cfg, i["latents"], i["timestep"], i["context"], y=i["y"], controlnet_sd=c_,
                                   control_signal_video_latents=i["control_signal_video_latents"],
                                   controlnet_num_layers=1).cpu())
        del s_, c_, i
        _free()
    ref32, refbf = outs
    pc = _prod_cfg(cfg)
    bf = {k: v.to("cuda", torch.bfloat16) for k, v in inp.items()}
    out = model_fn_wan_video(dit=WanModelB200(pc, sd), controlnet=ControlNetB200(pc, csd, 1), latents=bf["latents"],
                             timestep=bf["timestep"], context=bf["context"], y=bf["y"],
                             control_signal_video_latents=bf["control_signal_video_latents"]).cpu()
    assert out.shape == (1, 16, 21, 90, 160)
    e_ours, e_ref = O.rel_l2(out, ref32), O.rel_l2(refbf, ref32)
    print(f"75,600 tokens (81x720x1280), 2+1 blocks: relL2 ours-vs-fp32 {e_ours:.3e} ref_bf16-vs-fp32 {e_ref:.3e} "
          f"ours-vs-ref_bf16 {O.rel_l2(out, refbf):.3e}")
    _free()
    assert _bound(e_ours, e_ref), (e_ours, e_ref)


def test_reference_modules_through_model_fn():
    """INTEGRATION.md recipe: `pipe.model_fn = model_fn_wan_video` with the pipeline's own nn.Modules (stand-ins with the
    reference's attribute tree and state_dict keys, oracle/ref_standins.py) living on the GPU in bf16.  They are
    converted on first use, on the module's device; the result equals the explicit-conversion path bit for bit;
    loading ControlNet weights afterwards (load_controlnet_weights) is picked up."""
    from goal_force_b200 import wan_dit as W
    from oracle import ref_standins as S
    cfg = O.DiTConfig(**{**O.WAN22_I2V_A14B.__dict__, "num_layers": 2})
    sd = O.random_state_dict(cfg, seed=2)
    csd = O.random_controlnet_state_dict(cfg, 1, seed=3)
    inp = {k: v.to("cuda", torch.bfloat16) for k, v in O.synthetic_inputs(cfg, 4, 40, 52, seed=4, timestep=990.0).items()}
    dev = torch.device("cuda", torch.cuda.current_device())
    dit_mod = S.wan_standin_from_cfg(cfg, sd, device=dev)
    cn_mod = S.controlnet_standin_from_cfg(cfg, 1, device=dev)           # zero-convs still zero: untrained ControlNet
    kw = dict(latents=inp["latents"], timestep=inp["timestep"], context=inp["context"], y=inp["y"],
              control_signal_video_latents=inp["control_signal_video_latents"], height=480, width=832, seed=0,
              tiled=True, cfg_scale=5.0, motion_controller=None, vace=None)
    base = W.model_fn_wan_video(dit=dit_mod, controlnet=cn_mod, **kw)   # converted lazily; no-op ControlNet skipped
    pc = _prod_cfg(cfg)
    dit = W.WanModelB200(pc, sd)
    assert torch.equal(base, W.model_fn_wan_video(dit=dit, **kw))
    conv = W._CONVERTED[id(dit_mod)][2]
    assert conv.device == dev and conv.cfg == pc
    cn_mod.load_state_dict({k: v.to(dev, torch.bfloat16) for k, v in csd.items()})      # load_controlnet_weights
    got = W.model_fn_wan_video(dit=dit_mod, controlnet=cn_mod, **kw)
    want = W.model_fn_wan_video(dit=dit, controlnet=W.ControlNetB200(pc, csd, 1), **kw)
    assert torch.equal(got, want) and not torch.equal(got, base)
    # and against the oracle
    s32 = {k: v.to("cuda") for k, v in sd.items()}
    c32 = {k: v.to("cuda") for k, v in csd.items()}
    i32 = {k: v.float() for k, v in inp.items()}
    with torch.no_grad():
        ref32 = O.model_fn(s32, cfg, i32["latents"], i32["timestep"], i32["context"], y=i32["y"], controlnet_sd=c32,
                           control_signal_video_latents=i32["control_signal_video_latents"], controlnet_num_layers=1)
    assert O.rel_l2(got, ref32) <= 1e-2
