"""Whole-forward parity on the GPU: goal_force_b200.model_fn_wan_video (CUDA kernels through the C ABI) against
  * the committed reference vectors (tests/golden/dit_forward.pt, produced by the unmodified reference), and
  * the oracle run on the same device in fp32 and in bf16 (oracle/wan_dit_oracle.py, bit-identical to the reference).

Tolerance (BASELINE.json: per-step bf16 output relative L2 <= 1e-2 against the reference forward; SURVEY F15: the
reference's own bf16 forward is 1.55e-2 away from its fp32 forward at config 1, so a bf16 implementation cannot be
held to 1e-2 against fp32):   relL2(ours, ref_fp32) <= max(1e-2, 1.0 * relL2(ref_bf16, ref_fp32)).
"""
import pytest
import torch

from oracle import wan_dit_oracle as O

pytestmark = pytest.mark.gpu

TOL_ABS = 1e-2


def _cfg(d):
    d = dict(d)
    d["patch_size"] = tuple(d["patch_size"])
    return O.DiTConfig(**d)


def _prod_cfg(cfg):
    from goal_force_b200.wan_dit import DiTConfig
    return DiTConfig(**cfg.__dict__)


@pytest.fixture(scope="module", autouse=True)
def _no_tf32(lib):
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    yield
    torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old


def _oracle_pair(cfg, sd32, inp32, csd32=None, n_cn=0):
    """oracle on the GPU in fp32 and in bf16 (== what eager reference PyTorch computes on this device)."""
    outs = []
    for dt in (torch.float32, torch.bfloat16):
        sd = {k: v.to("cuda", dt) for k, v in sd32.items()}
        inp = {k: v.to("cuda", dt) for k, v in inp32.items()}
        kw = {}
        if csd32 is not None:
            kw = dict(controlnet_sd={k: v.to("cuda", dt) for k, v in csd32.items()},
                      control_signal_video_latents=inp["control_signal_video_latents"], controlnet_num_layers=n_cn)
        with torch.no_grad():
            outs.append(O.model_fn(sd, cfg, inp["latents"], inp["timestep"], inp["context"], y=inp.get("y"), **kw))
        del sd
    return outs


def _ours(cfg, sd32, inp32, csd32=None, n_cn=0):
    from goal_force_b200.wan_dit import ControlNetB200, WanModelB200, model_fn_wan_video
    dit = WanModelB200(_prod_cfg(cfg), sd32, device="cuda")
    bf = {k: v.to("cuda", torch.bfloat16) for k, v in inp32.items()}
    kw = {}
    if csd32 is not None:
        kw = dict(controlnet=ControlNetB200(_prod_cfg(cfg), csd32, n_cn, device="cuda"),
                  control_signal_video_latents=bf["control_signal_video_latents"])
    out = model_fn_wan_video(dit=dit, latents=bf["latents"], timestep=bf["timestep"], context=bf["context"],
                             y=bf.get("y"), height=480, width=832, seed=0, tiled=True, cfg_scale=5.0, **kw)
    return out, dit


def _check(out, ref32, refbf):
    e_ours, e_ref = O.rel_l2(out, ref32), O.rel_l2(refbf, ref32)
    e_bf = O.rel_l2(out, refbf)
    print(f"relL2 ours-vs-fp32 {e_ours:.3e}  ref_bf16-vs-fp32 {e_ref:.3e}  ours-vs-ref_bf16 {e_bf:.3e}")
    assert not torch.isnan(out).any()
    assert e_ours <= max(TOL_ABS, 1.0 * e_ref), (e_ours, e_ref)
    return e_ours, e_ref


@pytest.mark.parametrize("name", ["tiny_i2v", "tiny_t2v"])
def test_forward_matches_reference_golden(golden_dir, name):
    g = torch.load(golden_dir / "dit_forward.pt", weights_only=False)[name]
    cfg = _cfg(g["cfg"])
    sd = O.random_state_dict(cfg, seed=g["weight_seed"])
    inp = O.synthetic_inputs(cfg, *g["shape"], seed=g["input_seed"], ctx_len=g["ctx_len"], ctx_valid=g["ctx_valid"],
                             timestep=g["timestep"])
    out, _ = _ours(cfg, sd, inp)
    assert out.shape == g["out_fp32"].shape and out.dtype == torch.bfloat16
    _check(out.cpu(), g["out_fp32"], g["out_bf16"])


def test_forward_controlnet_matches_reference_golden(golden_dir):
    g = torch.load(golden_dir / "dit_forward.pt", weights_only=False)["a14b_slice_controlnet"]
    cfg = _cfg(g["cfg"])
    sd = O.random_state_dict(cfg, seed=g["weight_seed"])
    csd = O.random_controlnet_state_dict(cfg, 1, seed=g["controlnet_seed"])
    inp = O.synthetic_inputs(cfg, *g["shape"], seed=g["input_seed"], ctx_len=g["ctx_len"], ctx_valid=g["ctx_valid"],
                             timestep=g["timestep"])
    ref32, refbf = _oracle_pair(cfg, sd, inp, csd, 1)
    assert O.rel_l2(ref32, g["out_fp32"]) < 1e-4           # GPU oracle agrees with the CPU reference vector
    out, dit = _ours(cfg, sd, inp, csd, 1)
    _check(out, g["out_fp32"], refbf)
    # the ControlNet branch must be visible, and a zero-conv ControlNet must be an exact no-op (design invariant)
    from goal_force_b200.wan_dit import ControlNetB200, model_fn_wan_video
    bf = {k: v.to("cuda", torch.bfloat16) for k, v in inp.items()}
    kw = dict(dit=dit, latents=bf["latents"], timestep=bf["timestep"], context=bf["context"], y=bf["y"])
    base = model_fn_wan_video(**kw)
    assert O.rel_l2(out, base) > 1e-3
    _, base_bf = _oracle_pair(cfg, sd, inp)
    _check(base, g["out_base_fp32"], base_bf)
    czero = O.random_controlnet_state_dict(cfg, 1, seed=g["controlnet_seed"], zero_convs=True)
    cn0 = ControlNetB200(_prod_cfg(cfg), czero, 1, device="cuda")
    noop = model_fn_wan_video(controlnet=cn0, control_signal_video_latents=bf["control_signal_video_latents"], **kw)
    assert torch.equal(noop, base)


def test_forward_config1_shape_vs_oracle():
    """BASELINE.json configs[0]: Wan2.1-T2V-1.3B-shape DiT (random init, all 30 blocks, dim 1536), one denoise step at
    17 frames 240x416 (L = 1950), against the oracle in fp32 and bf16 on the same device."""
    cfg = O.WAN21_T2V_1_3B
    sd = O.random_state_dict(cfg, seed=0)
    inp = O.synthetic_inputs(cfg, 5, 30, 52, seed=1, timestep=900.0)
    ref32, refbf = _oracle_pair(cfg, sd, inp)
    out, dit = _ours(cfg, sd, inp)
    _check(out, ref32, refbf)
    # WanModel.forward signature == model_fn, and the call is deterministic
    bf = {k: v.to("cuda", torch.bfloat16) for k, v in inp.items()}
    fwd = dit(bf["latents"], bf["timestep"], bf["context"])
    assert torch.equal(fwd, out)


def test_a14b_width_two_blocks_long_sequence():
    """A14B widths (d 5120, 40 heads, ffn 13824), 2 blocks + 1 ControlNet block, L = 2080 tokens (ragged tiles:
    2080 = 16.25 x 128 = 26 x 80), against the oracle on the same device."""
    cfg = O.DiTConfig(**{**O.WAN22_I2V_A14B.__dict__, "num_layers": 2})
    sd = O.random_state_dict(cfg, seed=2)
    csd = O.random_controlnet_state_dict(cfg, 1, seed=3)
    inp = O.synthetic_inputs(cfg, 4, 40, 52, seed=4, timestep=990.0)
    ref32, refbf = _oracle_pair(cfg, sd, inp, csd, 1)
    out, _ = _ours(cfg, sd, inp, csd, 1)
    _check(out, ref32, refbf)


def test_sampler_matches_oracle_loop():
    """4-step two-expert CFG sampling loop (expert switch at t < 875, CFG 5.0, shift 5.0) against the same loop
    written with the oracle forward in bf16; final-latent cosine >= 0.999 (BASELINE.json criterion)."""
    from goal_force_b200.pipeline import GoalForceDenoiser
    from goal_force_b200.wan_dit import WanModelB200
    from goal_force_b200.scheduler import FlowMatchScheduler
    cfg = O.DiTConfig(dim=256, in_dim=36, ffn_dim=512, out_dim=16, text_dim=64, freq_dim=256, eps=1e-6, num_heads=2,
                      num_layers=2)
    sds = [O.random_state_dict(cfg, seed=s) for s in (10, 11)]
    inp = O.synthetic_inputs(cfg, 3, 8, 12, seed=12, ctx_len=32, ctx_valid=8)
    g = torch.Generator("cpu").manual_seed(13)
    ctx_n = torch.randn(1, 32, cfg.text_dim, generator=g)
    bf = {k: v.to("cuda", torch.bfloat16) for k, v in inp.items()}
    ctx_n = ctx_n.to("cuda", torch.bfloat16)
    den = GoalForceDenoiser(WanModelB200(_prod_cfg(cfg), sds[0]), WanModelB200(_prod_cfg(cfg), sds[1]))
    got = den(bf["latents"], bf["context"], ctx_n, y=bf["y"], num_inference_steps=4, cfg_scale=5.0, sigma_shift=5.0)
    # oracle loop (wan_video_new.py:697-721) with eager bf16 torch ops
    sch = FlowMatchScheduler()
    sch.set_timesteps(4, shift=5.0)
    sdb = [{k: v.to("cuda", torch.bfloat16) for k, v in sd.items()} for sd in sds]
    lat = bf["latents"]
    used = []
    with torch.no_grad():
        for i, t in enumerate(sch.timesteps):
            e = 1 if float(t) < 875 else 0
            used.append(e)
            ts = t.unsqueeze(0).to("cuda", torch.bfloat16)
            p = O.model_fn(sdb[e], cfg, lat, ts, bf["context"], y=bf["y"])
            n = O.model_fn(sdb[e], cfg, lat, ts, ctx_n, y=bf["y"])
            pred = n + 5.0 * (p - n)
            s0, s1 = sch.sigma_pair(t)
            lat = lat + pred * (s1 - s0)
    assert used == [0, 0, 1, 1]              # t = 1000, 937.5 -> high-noise expert; 833, 625 -> low-noise expert
    cos = O.cosine(got, lat)
    print("sampler cosine", cos, "relL2", O.rel_l2(got, lat))
    assert cos >= 0.999


def test_sampler_40_steps_two_experts_controlnet_cosine():
    """BASELINE.json criterion: final latent after a fixed-seed 40-step sampling run (two experts switching at
    t < 875 -> 17 high-noise + 23 low-noise steps, CFG 5.0, shift 5.0, goal-force ControlNet on the high-noise expert,
    an all-zero -- i.e. skipped -- ControlNet on the low-noise expert as in the shipped inference script) has cosine
    similarity >= 0.999 against the same loop run with the oracle forward in bf16 on this device."""
    from goal_force_b200.pipeline import GoalForceDenoiser, generate_noise
    from goal_force_b200.scheduler import FlowMatchScheduler
    from goal_force_b200.wan_dit import ControlNetB200, WanModelB200
    cfg = O.DiTConfig(dim=1536, in_dim=36, ffn_dim=4096, out_dim=16, text_dim=256, freq_dim=256, eps=1e-6,
                      num_heads=12, num_layers=3)
    sds = [O.random_state_dict(cfg, seed=s) for s in (20, 21)]
    csd = O.random_controlnet_state_dict(cfg, 2, seed=22)
    csd0 = O.random_controlnet_state_dict(cfg, 2, seed=23, zero_convs=True)
    inp = O.synthetic_inputs(cfg, 3, 16, 24, seed=24, ctx_len=64, ctx_valid=16)
    bf = {k: v.to("cuda", torch.bfloat16) for k, v in inp.items()}
    g = torch.Generator("cpu").manual_seed(25)
    ctx_n = torch.randn(1, 64, cfg.text_dim, generator=g).to("cuda", torch.bfloat16)
    noise = generate_noise(tuple(inp["latents"].shape), seed=5)
    pc = _prod_cfg(cfg)
    den = GoalForceDenoiser(WanModelB200(pc, sds[0]), WanModelB200(pc, sds[1]), ControlNetB200(pc, csd, 2),
                            ControlNetB200(pc, csd0, 2))
    assert den.controlnet2.is_noop and not den.controlnet.is_noop
    got = den(noise, bf["context"], ctx_n, y=bf["y"], control_latents=bf["control_signal_video_latents"],
              num_inference_steps=40, cfg_scale=5.0, sigma_shift=5.0)
    sch = FlowMatchScheduler()
    sch.set_timesteps(40, shift=5.0)
    sdb = [{k: v.to("cuda", torch.bfloat16) for k, v in sd.items()} for sd in sds]
    cb = {k: v.to("cuda", torch.bfloat16) for k, v in csd.items()}
    lat, used = noise, []
    with torch.no_grad():
        for t in sch.timesteps:
            e = 1 if float(t) < 875 else 0
            used.append(e)
            ts = t.unsqueeze(0).to("cuda", torch.bfloat16)
            kw = dict(y=bf["y"])
            if e == 0:      # controlnet2's zero-convs are all zero: the reference result equals the plain forward
                kw.update(controlnet_sd=cb, control_signal_video_latents=bf["control_signal_video_latents"],
                          controlnet_num_layers=2)
            p = O.model_fn(sdb[e], cfg, lat, ts, bf["context"], **kw)
            n = O.model_fn(sdb[e], cfg, lat, ts, ctx_n, **kw)
            pred = n + 5.0 * (p - n)
            s0, s1 = sch.sigma_pair(t)
            lat = lat + pred * (s1 - s0)
    assert used.count(0) == 17 and used.count(1) == 23
    cos = O.cosine(got, lat)
    print("40-step sampler cosine", cos, "relL2", O.rel_l2(got, lat))
    assert cos >= 0.999


def test_models_from_safetensors_equal_models_from_state_dict(tmp_path):
    """goal_force_b200.checkpoint: sharded expert + 'pipe.controlnet.'-prefixed ControlNet checkpoint give the same
    forward, bit for bit, as the in-memory state dicts."""
    from safetensors.torch import save_file
    from goal_force_b200.checkpoint import load_controlnet, load_dit
    from goal_force_b200.wan_dit import ControlNetB200, WanModelB200, model_fn_wan_video
    cfg = O.DiTConfig(dim=256, in_dim=36, ffn_dim=512, out_dim=16, text_dim=64, freq_dim=256, eps=1e-6, num_heads=2,
                      num_layers=2)
    sd = {k: v.bfloat16().contiguous() for k, v in O.random_state_dict(cfg, seed=30).items()}
    csd = {k: v.bfloat16().contiguous() for k, v in O.random_controlnet_state_dict(cfg, 1, seed=31).items()}
    keys = sorted(sd)
    save_file({k: sd[k] for k in keys[::2]}, str(tmp_path / "diffusion_pytorch_model-00001-of-00002.safetensors"))
    save_file({k: sd[k] for k in keys[1::2]}, str(tmp_path / "diffusion_pytorch_model-00002-of-00002.safetensors"))
    save_file({"pipe.controlnet." + k: v for k, v in csd.items()}, str(tmp_path / "step-3000.safetensors"))
    pc = _prod_cfg(cfg)
    dit_a, cn_a = WanModelB200(pc, sd), ControlNetB200(pc, csd, 1)
    dit_b = load_dit(sorted(tmp_path.glob("diffusion_pytorch_model-*.safetensors")), pc)
    cn_b = load_controlnet(tmp_path / "step-3000.safetensors", pc, 1)
    inp = {k: v.to("cuda", torch.bfloat16) for k, v in O.synthetic_inputs(cfg, 2, 8, 12, seed=32, ctx_len=32, ctx_valid=8).items()}
    kw = dict(latents=inp["latents"], timestep=inp["timestep"], context=inp["context"], y=inp["y"],
              control_signal_video_latents=inp["control_signal_video_latents"])
    assert torch.equal(model_fn_wan_video(dit=dit_a, controlnet=cn_a, **kw), model_fn_wan_video(dit=dit_b, controlnet=cn_b, **kw))


def test_cuda_graph_replay_equals_eager():
    """GraphedModelFn: one capture per (expert, shapes), every tensor input static, step-invariants recomputed inside
    the graph; replays with new latents / timestep / prompt / control latents equal the eager forward bit for bit, and
    the descriptor cache of the context is hit on every launch after the first forward."""
    from goal_force_b200 import capi
    from goal_force_b200.graph import GraphedModelFn
    from goal_force_b200.wan_dit import ControlNetB200, WanModelB200, model_fn_wan_video
    cfg = O.DiTConfig(dim=1536, in_dim=36, ffn_dim=4096, out_dim=16, text_dim=256, freq_dim=256, eps=1e-6,
                      num_heads=12, num_layers=2)
    pc = _prod_cfg(cfg)
    dit = WanModelB200(pc, O.random_state_dict(cfg, seed=50))
    cn = ControlNetB200(pc, O.random_controlnet_state_dict(cfg, 1, seed=51), 1)
    gfn = GraphedModelFn()
    before = capi.ctx_stats()
    for k in range(3):
        inp = O.synthetic_inputs(cfg, 3, 16, 24, seed=60 + k, ctx_len=64, ctx_valid=16, timestep=990.0 - 100 * k)
        bf = {n: v.to("cuda", torch.bfloat16) for n, v in inp.items()}
        kw = dict(dit=dit, controlnet=cn, latents=bf["latents"], timestep=bf["timestep"], context=bf["context"],
                  y=bf["y"], control_signal_video_latents=bf["control_signal_video_latents"])
        want = model_fn_wan_video(**kw)
        got = gfn(**kw)
        assert torch.equal(got, want), k
    assert gfn.captures == 1 and gfn.replays == 3
    after = capi.ctx_stats()
    assert after["tmap_hits"] > before["tmap_hits"] and after["tmap_entries"] > 0


def test_controlnet_on_second_stream_is_bit_identical():
    """the ControlNet branch never reads the noisy latents, so it may run on its own stream next to the trunk
    (per-block events hand the states over).  Same kernels, same operands: bit-identical, also for the strided inject
    and across repeated calls that reuse the branch's workspace."""
    from goal_force_b200.wan_dit import ControlNetB200, WanModelB200, model_fn_wan_video
    cfg = O.DiTConfig(dim=1536, in_dim=36, ffn_dim=4096, out_dim=16, text_dim=256, freq_dim=256, eps=1e-6,
                      num_heads=12, num_layers=4)
    pc = _prod_cfg(cfg)
    dit = WanModelB200(pc, O.random_state_dict(cfg, seed=80))
    csd = O.random_controlnet_state_dict(cfg, 2, seed=81)
    for stride in (None, 2):
        cn = ControlNetB200(pc, csd, 2, stride=stride)
        outs = {}
        for k in range(2):
            inp = O.synthetic_inputs(cfg, 3, 16, 24, seed=90 + k, ctx_len=64, ctx_valid=16, timestep=950.0)
            bf = {n: v.to("cuda", torch.bfloat16) for n, v in inp.items()}
            kw = dict(dit=dit, controlnet=cn, latents=bf["latents"], timestep=bf["timestep"], context=bf["context"],
                      y=bf["y"], control_signal_video_latents=bf["control_signal_video_latents"])
            a = model_fn_wan_video(controlnet_stream=False, **kw)
            b = model_fn_wan_video(controlnet_stream=True, **kw)
            assert torch.equal(a, b), (stride, k)
            outs[k] = a
        assert not torch.equal(outs[0], outs[1])
