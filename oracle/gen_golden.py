"""Generate tests/golden/* by running the UNMODIFIED reference from /root/reference on CPU -- TEST INFRASTRUCTURE ONLY.

    python -m oracle.gen_golden            # writes tests/golden/{dit_*.pt, scheduler.json, control_channels.json}

The fixtures hold inputs and reference outputs only; weights are re-created from a seed by
oracle.wan_dit_oracle.random_state_dict (a pure torch.Generator recipe), so the files stay small.
Every vector here was produced by reference code: WanModel / model_fn_wan_video / ControlNet
(src/goal_force/wan_video_new.py, diffsynth/models/wan_video_dit.py), FlowMatchScheduler
(diffsynth/schedulers/flow_match.py) and ControlSignalDataset_Balls (src/goal_force/unified_dataset.py).
"""
from __future__ import annotations

import glob
import json
import os
import shutil
import tempfile
from pathlib import Path

import numpy as np
import torch

from . import control_channels_oracle as CC
from . import ref_shim
from . import wan_dit_oracle as O

GOLDEN = Path(__file__).resolve().parent.parent / "tests" / "golden"

TINY = O.DiTConfig(dim=256, in_dim=36, ffn_dim=512, out_dim=16, text_dim=64, freq_dim=256, eps=1e-6, num_heads=2,
                   num_layers=2)
TINY_T2V = O.DiTConfig(dim=256, in_dim=16, ffn_dim=512, out_dim=16, text_dim=64, freq_dim=256, eps=1e-6, num_heads=2,
                       num_layers=3)
# ControlNet_DiT hard-codes the A14B widths (src/goal_force/wan_video_new.py:56-60), so the ControlNet golden has to
# be A14B-wide; two trunk blocks + one ControlNet block keep it runnable on CPU.
A14B_SLICE = O.DiTConfig(dim=5120, in_dim=36, ffn_dim=13824, out_dim=16, text_dim=4096, freq_dim=256, eps=1e-6,
                         num_heads=40, num_layers=2)


def _ref_model(ns, cfg, sd):
    m = ns.WanModel(**ref_shim.cfg_kwargs(cfg)).eval()
    m.load_state_dict(sd, strict=True)
    return m


def gen_dit(ns):
    out = {}
    with torch.no_grad():
        # 1) tiny I2V-shaped model, model_fn without ControlNet, fp32 and bf16, plus WanModel.forward equivalence (F8)
        for name, cfg, shape in (("tiny_i2v", TINY, (3, 8, 12)), ("tiny_t2v", TINY_T2V, (2, 6, 10))):
            sd = O.random_state_dict(cfg, seed=0)
            inp = O.synthetic_inputs(cfg, *shape, seed=1, ctx_len=32, ctx_valid=8, timestep=937.0)
            m = _ref_model(ns, cfg, sd)
            kw = dict(latents=inp["latents"], timestep=inp["timestep"], context=inp["context"])
            if "y" in inp:
                kw["y"] = inp["y"]
            ref = ns.model_fn_wan_video(dit=m, **kw)
            x_in = torch.cat([inp["latents"], inp["y"]], 1) if "y" in inp else inp["latents"]
            fwd = m(x_in, inp["timestep"], inp["context"])
            assert torch.equal(ref, fwd), "model_fn != WanModel.forward(cat[x,y]) (SURVEY F8)"
            mb = _ref_model(ns, cfg, sd).to(torch.bfloat16)
            kwb = {k: v.to(torch.bfloat16) for k, v in kw.items()}
            ref_bf16 = ns.model_fn_wan_video(dit=mb, **kwb)
            out[name] = dict(cfg=cfg.__dict__, weight_seed=0, input_seed=1, shape=shape, ctx_len=32, ctx_valid=8,
                             timestep=937.0, out_fp32=ref, out_bf16=ref_bf16)
            print(name, "fp32-vs-bf16 relL2", O.rel_l2(ref_bf16, ref))
        # 2) A14B-wide slice with the reference ControlNet (non-zero zero-convs) and the zero-conv no-op invariant
        cfg = A14B_SLICE
        sd = O.random_state_dict(cfg, seed=0)
        csd = O.random_controlnet_state_dict(cfg, 1, seed=1)
        inp = O.synthetic_inputs(cfg, 2, 8, 12, seed=1, ctx_len=32, ctx_valid=8, timestep=937.0)
        m = _ref_model(ns, cfg, sd)
        cn = ns.ControlNet(1, stride=None, torch_dtype=torch.float32).eval()
        cn.load_state_dict(csd, strict=True)
        kw = dict(latents=inp["latents"], timestep=inp["timestep"], context=inp["context"], y=inp["y"])
        base = ns.model_fn_wan_video(dit=m, **kw)
        with_cn = ns.model_fn_wan_video(dit=m, controlnet=cn,
                                        control_signal_video_latents=inp["control_signal_video_latents"], **kw)
        czero = O.random_controlnet_state_dict(cfg, 1, seed=1, zero_convs=True)
        cn.load_state_dict(czero, strict=True)
        noop = ns.model_fn_wan_video(dit=m, controlnet=cn,
                                     control_signal_video_latents=inp["control_signal_video_latents"], **kw)
        assert torch.equal(noop, base), "zero-conv ControlNet is not a no-op"
        out["a14b_slice_controlnet"] = dict(cfg=cfg.__dict__, weight_seed=0, controlnet_seed=1, input_seed=1,
                                            shape=(2, 8, 12), ctx_len=32, ctx_valid=8, timestep=937.0,
                                            out_fp32=with_cn, out_base_fp32=base)
        print("a14b slice: controlnet effect relL2", O.rel_l2(with_cn, base))
    torch.save(out, GOLDEN / "dit_forward.pt")


def gen_scheduler(ns):
    res = {}
    for steps, shift in ((40, 5.0), (50, 5.0), (4, 5.0)):
        s = ns.FlowMatchScheduler(shift=5, sigma_min=0.0, extra_one_step=True)   # as WanVideoPipeline.__init__ (:129)
        s.set_timesteps(steps, denoising_strength=1.0, shift=shift)
        n_high = int((s.timesteps >= 875).sum())
        # one Euler step on a fixed bf16 sample through the reference scheduler
        g = torch.Generator("cpu").manual_seed(3)
        sample = torch.randn(2, 3, 4, generator=g).bfloat16()
        pred = torch.randn(2, 3, 4, generator=g).bfloat16()
        stepped = [s.step(pred, s.timesteps[i], sample).float().tolist() for i in (0, steps // 2, steps - 1)]
        res[f"{steps}_{shift}"] = dict(sigmas=s.sigmas.tolist(), timesteps=s.timesteps.tolist(), n_high_noise=n_high,
                                       sample=sample.float().tolist(), pred=pred.float().tolist(), stepped=stepped)
    (GOLDEN / "scheduler.json").write_text(json.dumps(res))


def _dataset_overrides(ds):
    # scripts/inference/inference_goal_force.py:137-144
    ds.min_mass, ds.max_mass = 1.0, 4.0
    ds.min_force, ds.max_force = 30.0, 400.0
    ds.min_indirect_force, ds.max_indirect_force = ds.min_force, ds.max_force


def gen_control_channels():
    import pandas
    ud = ref_shim.load_dataset_module()
    csvs = sorted(glob.glob(os.path.join(ref_shim.REFERENCE_ROOT, "datasets/examples/*/*.csv")))
    csvs = [c for c in csvs if "canny" not in c]
    res = {"goal_force": {}, "direct_force": {}, "rows": {}}
    tmp = Path(tempfile.mkdtemp(prefix="gf_golden_"))
    try:
        for csv in csvs:
            name = os.path.basename(csv).split("_obj")[0]
            ds = ud.ControlSignalDataset_Balls(base_path=os.path.dirname(csv), metadata_path=csv,
                                               is_validation_dataset=True, num_frames=81, height=480, width=832)
            _dataset_overrides(ds)
            np.random.seed(0)
            cv = ds[0]["control_video"]
            res["goal_force"][name] = CC.digest(cv)
            row = pandas.read_csv(csv).iloc[0].to_dict()
            res["rows"][name] = {k: (v.item() if hasattr(v, "item") else v) for k, v in row.items() if k != "caption"}
            # direct-force variant: give the projectile a force + mass, drop the goal force
            df = pandas.read_csv(csv)
            for col in ("projectile_force_magnitude", "projectile_force_angle", "projectile_mass",
                        "target_indirect_force_magnitude"):
                df[col] = df[col].astype(float)
            df.loc[0, "projectile_force_magnitude"] = 250.0
            df.loc[0, "projectile_force_angle"] = 37.0
            df.loc[0, "projectile_mass"] = 2.5
            df.loc[0, "target_indirect_force_magnitude"] = -1
            d = tmp / name
            (d / "images").mkdir(parents=True)
            img = df.loc[0, "image"]
            os.symlink(os.path.join(os.path.dirname(csv), "images", img), d / "images" / img)
            df.to_csv(d / "row.csv", index=False)
            ds2 = ud.ControlSignalDataset_Balls(base_path=str(d), metadata_path=str(d / "row.csv"),
                                                is_validation_dataset=True, num_frames=81, height=480, width=832)
            _dataset_overrides(ds2)
            np.random.seed(0)
            cv2 = ds2[0]["control_video"]
            res["direct_force"][name] = CC.digest(cv2)
            print(name, res["goal_force"][name], res["direct_force"][name], tuple(cv.shape))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    (GOLDEN / "control_channels.json").write_text(json.dumps(res, indent=1))


class _FakeImage:
    """Stands in for the PIL image the reference unit receives: .resize((w, h)) returns itself."""

    def __init__(self, seed):
        self.seed = seed

    def resize(self, size):
        self.size = size
        return self


class _FakePipe:
    """The attributes WanVideoUnit_ImageEmbedderVAE.process touches (src/goal_force/wan_video_new.py:894-917); the VAE
    is replaced by a seeded random latent generator -- the mask / concat bookkeeping under test is the reference's."""

    def __init__(self):
        import types
        self.device, self.torch_dtype = "cpu", torch.bfloat16
        self.dit = types.SimpleNamespace(require_vae_embedding=True)
        self.vae = types.SimpleNamespace(encode=self._encode)

    def load_models_to_device(self, names):
        pass

    def preprocess_image(self, img):
        w, h = img.size
        g = torch.Generator("cpu").manual_seed(img.seed)
        return torch.rand(1, 3, h, w, generator=g) * 2 - 1

    def _encode(self, videos, device=None, tiled=False, tile_size=None, tile_stride=None):
        v = videos[0]                                  # (3, num_frames, H, W)
        self.vae_input_digest = CC.digest(v.to(torch.bfloat16))
        g = torch.Generator("cpu").manual_seed(77)
        return [torch.randn(16, (v.shape[1] - 1) // 4 + 1, v.shape[2] // 8, v.shape[3] // 8, generator=g)]


def vae_latents_for_mask_golden(num_frames, height, width):
    g = torch.Generator("cpu").manual_seed(77)
    return torch.randn(16, (num_frames - 1) // 4 + 1, height // 8, width // 8, generator=g)


def gen_mask(ns):
    """a21: y = cat(mask, vae_latents) exactly as the reference unit builds it (incl. the end-image variant)."""
    unit = ns.pipe_mod.WanVideoUnit_ImageEmbedderVAE()
    out = {}
    for name, (nf, h, w, end) in {"small": (81, 64, 96, False), "small_end": (81, 64, 96, True),
                                  "odd_frames": (17, 32, 48, False), "full": (81, 480, 832, False)}.items():
        pipe = _FakePipe()
        r = unit.process(pipe, _FakeImage(1), _FakeImage(2) if end else None, nf, h, w, True, (30, 52), (15, 26))
        y = r["y"]
        assert y.dtype == torch.bfloat16 and y.shape == (1, 20, (nf - 1) // 4 + 1, h // 8, w // 8)
        ent = dict(num_frames=nf, height=h, width=w, end_image=end, digest=CC.digest(y))
        if name != "full":
            ent["y"] = y
        out[name] = ent
        print("mask", name, tuple(y.shape), ent["digest"])
    torch.save(out, GOLDEN / "image_condition.pt")


def gen_umt5():
    """N4: the reference WanTextEncoder (eval mode) on seeded weights, fp32 and bf16, with a padding mask; plus the
    prompter's zeroing of the positions past the prompt length (wan_prompter.py:105-108)."""
    from . import umt5_oracle as U
    te = ref_shim.load_module("diffsynth.models.wan_video_text_encoder")
    cfgs = {"tiny": dict(vocab=97, dim=256, dim_attn=256, dim_ffn=512, num_heads=4, num_layers=2, num_buckets=32)}
    out = {}
    with torch.no_grad():
        for name, c in cfgs.items():
            sd = U.random_state_dict(seed=0, **c)
            ids, mask = U.synthetic_prompt(c["vocab"], 2, 48, (48, 19), seed=1)
            m = te.WanTextEncoder(shared_pos=False, dropout=0.1, **c).eval()
            m.load_state_dict(sd, strict=True)
            ref32 = m(ids, mask)
            mb = te.WanTextEncoder(shared_pos=False, dropout=0.1, **c).eval()
            mb.load_state_dict(sd, strict=True)
            mb = mb.to(torch.bfloat16)
            refbf = mb(ids, mask)
            assert torch.equal(ref32, U.encoder(sd, ids, mask, num_heads=c["num_heads"], num_layers=c["num_layers"],
                                                num_buckets=c["num_buckets"]))
            out[name] = dict(cfg=c, weight_seed=0, prompt_seed=1, batch=2, L=48, valid=(48, 19), out_fp32=ref32,
                             out_bf16=refbf, buckets=m.blocks[0].pos_embedding._relative_position_bucket(
                                 torch.arange(48).unsqueeze(0) - torch.arange(48).unsqueeze(1)).to(torch.int32))
            print("umt5", name, tuple(ref32.shape), "fp32-vs-bf16 relL2", O.rel_l2(refbf, ref32))
    torch.save(out, GOLDEN / "umt5.pt")


def gen_vae():
    """N2: the reference Wan2.1 video VAE (VideoVAE_ chunked encode / decode and WanVideoVAE's tiled wrappers) on
    seeded weights at a CPU-sized width (dim 32 instead of 96; same topology), fp32.  Also asserts that the full-clip
    formulation of oracle/wan_vae_oracle.py equals the reference's chunk-by-chunk walk."""
    from . import wan_vae_oracle as V
    vae = ref_shim.load_module("diffsynth.models.wan_video_vae")
    dim = 32
    sd = V.random_state_dict(dim=dim, seed=0)
    wrap = vae.WanVideoVAE(z_dim=16)
    wrap.model = vae.VideoVAE_(dim=dim, z_dim=16).eval().requires_grad_(False)
    wrap.model.load_state_dict(sd, strict=True)
    g = torch.Generator().manual_seed(1)
    z = torch.randn(1, 16, 3, 4, 6, generator=g)
    video = torch.randn(1, 3, 9, 32, 48, generator=g).clamp_(-1, 1)
    z_big = torch.randn(1, 16, 2, 7, 9, generator=g)
    out = {"dim": dim, "weight_seed": 0, "z": z, "video": video, "z_big": z_big}
    with torch.no_grad():
        out["decode"] = wrap.model.decode(z, wrap.scale)
        out["encode"] = wrap.model.encode(video, wrap.scale)
        out["tiled_decode"] = wrap.tiled_decode(z_big, "cpu", (4, 5), (3, 3))
        out["tiled_encode"] = wrap.tiled_encode(out["tiled_decode"], "cpu", (32, 40), (24, 24))
        assert O.rel_l2(V.decode(sd, z, dim=dim), out["decode"]) < 1e-5
        assert O.rel_l2(V.encode(sd, video, dim=dim), out["encode"]) < 1e-5
        assert O.rel_l2(V.tiled_decode(sd, z_big, (4, 5), (3, 3), dim=dim), out["tiled_decode"]) < 1e-5
        assert O.rel_l2(V.tiled_encode(sd, out["tiled_decode"], (32, 40), (24, 24), dim=dim), out["tiled_encode"]) < 1e-5
        # the shipped width (dim 96: 96 / 192 / 384 channels) on a tiny clip, again from the reference's own chunk walk
        sd96 = V.random_state_dict(dim=96, seed=7)
        ref96 = vae.VideoVAE_(dim=96, z_dim=16).eval().requires_grad_(False)
        ref96.load_state_dict(sd96, strict=True)
        z96 = torch.randn(1, 16, 2, 3, 4, generator=g)
        out["z96"], out["weight_seed96"] = z96, 7
        out["decode96"] = ref96.decode(z96, wrap.scale)
        out["encode96"] = ref96.encode(out["decode96"].clamp(-1, 1), wrap.scale)
        assert O.rel_l2(V.decode(sd96, z96), out["decode96"]) < 1e-5
        assert O.rel_l2(V.encode(sd96, out["decode96"].clamp(-1, 1)), out["encode96"]) < 1e-5
    out = {k: (v.to(torch.float16) if isinstance(v, torch.Tensor) and k in ("decode", "tiled_decode", "decode96") else v)
           for k, v in out.items()}
    for k in ("decode", "encode", "tiled_decode", "tiled_encode", "decode96", "encode96"):
        print("vae", k, tuple(out[k].shape))
    torch.save(out, GOLDEN / "vae.pt")


def main():
    GOLDEN.mkdir(parents=True, exist_ok=True)
    torch.set_num_threads(os.cpu_count() or 1)
    ns = ref_shim.load()
    import sys
    what = sys.argv[1:] or ["scheduler", "dit", "control", "mask", "umt5", "vae"]
    if "scheduler" in what:
        gen_scheduler(ns)
    if "dit" in what:
        gen_dit(ns)
    if "control" in what:
        gen_control_channels()
    if "mask" in what:
        gen_mask(ns)
    if "umt5" in what:
        gen_umt5()
    if "vae" in what:
        gen_vae()


if __name__ == "__main__":
    main()
