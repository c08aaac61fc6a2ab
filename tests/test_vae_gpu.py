"""Wan video VAE on B200: the implicit-GEMM convolution and the streaming kernels against plain torch
fp32 on the same bf16 inputs, and the whole encoder / decoder / tiling against the oracle and the reference's golden
vectors (tests/golden/vae.pt).  Tolerance for whole-network comparisons: our distance to the fp32 oracle may not exceed
the distance of the oracle's own bf16 run (same weights, same input) to it -- stated per test."""
import pytest
import torch
import torch.nn.functional as F

from oracle import wan_dit_oracle as O
from oracle import wan_vae_oracle as V

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _no_tf32():
    a, b = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = a, b


def _bf(t):
    return t.to(torch.bfloat16)


def _cl(x):
    """(C, T, H, W) -> [T, H, W, C] contiguous"""
    return x.permute(1, 2, 3, 0).contiguous()


def _w2d(w, cinp):
    cout, cin = w.shape[:2]
    w = w.permute(0, 2, 3, 4, 1)
    if cinp != cin:
        w = F.pad(w, (0, cinp - cin))
    return w.reshape(cout, -1).contiguous()


CONV_CASES = [
    # name, Cin, Cout, (T, H, W), kernel, stride, pad(leading), (residual, fused norm, ncthw)
    ("causal333_96", 96, 96, (5, 20, 28), (3, 3, 3), (1, 1, 1), (2, 1, 1), (False, False, False)),
    ("causal333_16_to_32", 16, 32, (3, 9, 13), (3, 3, 3), (1, 1, 1), (2, 1, 1), (False, False, False)),
    ("causal333_8_to_96", 8, 96, (5, 16, 24), (3, 3, 3), (1, 1, 1), (2, 1, 1), (False, True, False)),
    ("causal333_192_res_norm", 192, 192, (4, 12, 20), (3, 3, 3), (1, 1, 1), (2, 1, 1), (True, True, False)),
    ("causal333_384_two_n_tiles", 384, 384, (3, 10, 14), (3, 3, 3), (1, 1, 1), (2, 1, 1), (True, False, False)),
    ("causal333_192_to_384", 192, 384, (2, 8, 8), (3, 3, 3), (1, 1, 1), (2, 1, 1), (False, False, False)),
    ("conv2d_33_pad1", 192, 96, (3, 16, 24), (1, 3, 3), (1, 1, 1), (0, 1, 1), (False, True, False)),
    ("conv2d_33_stride2", 96, 96, (3, 16, 24), (1, 3, 3), (1, 2, 2), (0, 0, 0), (False, False, False)),
    ("conv2d_33_stride2_odd_tiles", 64, 64, (2, 30, 52), (1, 3, 3), (1, 2, 2), (0, 0, 0), (False, True, False)),
    ("time311_causal_to_2c", 128, 256, (4, 6, 10), (3, 1, 1), (1, 1, 1), (2, 0, 0), (False, False, False)),
    ("time311_stride2", 192, 192, (9, 6, 10), (3, 1, 1), (2, 1, 1), (0, 0, 0), (False, False, False)),
    ("pointwise_16", 16, 16, (3, 7, 9), (1, 1, 1), (1, 1, 1), (0, 0, 0), (False, False, False)),
    ("head_96_to_3_ncthw", 96, 3, (5, 24, 40), (3, 3, 3), (1, 1, 1), (2, 1, 1), (False, False, True)),
    ("wide_frame_128", 96, 96, (2, 8, 300), (3, 3, 3), (1, 1, 1), (2, 1, 1), (False, False, False)),
]


@pytest.mark.parametrize("case", CONV_CASES, ids=[c[0] for c in CONV_CASES])
def test_conv3d_against_torch(capi, case):
    name, cin, cout, (T, H, W), kernel, stride, pad, (use_res, use_norm, ncthw) = case
    g = torch.Generator(device="cpu").manual_seed(sum(name.encode()))
    kt, kh, kw = kernel
    x = _bf(torch.randn(cin, T, H, W, generator=g)).cuda()
    w = _bf(torch.randn(cout, cin, kt, kh, kw, generator=g) * (cin * kt * kh * kw) ** -0.5).cuda()
    b = _bf(torch.randn(cout, generator=g) * 0.1).cuda()
    # reference: explicit zero padding (leading pad as given, trailing so that the output size is what the VAE uses)
    if stride[1] == 2:
        Ho, Wo = H // 2, W // 2
        xp = F.pad(x.float(), (0, 1, 0, 1, pad[0], 0))
    else:
        Ho, Wo = H, W
        xp = F.pad(x.float(), (pad[2], kw - 1 - pad[2], pad[1], kh - 1 - pad[1], pad[0], 0))
    ref = F.conv3d(xp.unsqueeze(0), w.float(), b.float(), stride=stride)[0]
    To = ref.shape[1]
    assert ref.shape == (cout, To, Ho, Wo)
    ref = _bf(ref).float()
    res = None
    if use_res:
        res = _bf(torch.randn(cout, To, Ho, Wo, generator=g)).cuda()
        ref = _bf(ref + res.float()).float()
    bias = torch.zeros((cout + 7) // 8 * 8, dtype=torch.bfloat16, device="cuda")
    bias[:cout] = b
    gamma = _bf(1.0 + 0.1 * torch.randn(cout, generator=g)).cuda() if use_norm else None
    y, yn = capi.conv3d_cl(_cl(x), _w2d(w, cin), bias, kernel=kernel, stride=stride, pad=pad, out_dims=(To, Ho, Wo),
                           residual=_cl(res) if use_res else None, gamma=gamma, silu=True, cout=cout, ncthw=ncthw)
    torch.cuda.synchronize()
    got = y.float() if ncthw else y.float().permute(3, 0, 1, 2)[:cout]
    err = O.rel_l2(got, ref)
    assert err < 3e-3, f"{name}: rel_l2 {err:.3e}, max abs {float((got - ref).abs().max()):.3e}"
    assert float((got - ref).abs().max()) <= 0.02 * float(ref.abs().max()) + 1e-2
    if use_norm:
        want = F.silu(V.rms_norm(ref.unsqueeze(0), gamma.float().view(-1, 1, 1, 1))[0])
        gotn = yn.float().permute(3, 0, 1, 2)[:cout]
        errn = O.rel_l2(gotn, want)
        assert errn < 6e-3, f"{name}: fused norm rel_l2 {errn:.3e}"
        sep = capi.vae_rmsnorm(y, gamma, silu=True)
        torch.cuda.synchronize()
        assert O.rel_l2(sep.float(), yn.float()) < 1e-3


def _sweep_cases():
    """Seeded random 3x3-window shapes for the halo kernels (narrow: Cout <= 128, wide: one accumulator, n-tiles) and the
    tap form (forced): odd frame sizes, single frames, frames narrower than a tile, Cin with a partial last k-block."""
    import random
    rng = random.Random(20260117)
    cases = []
    for i in range(24):
        cin = rng.choice([8, 16, 32, 64, 96, 160, 192])
        cout = rng.choice([3, 16, 32, 48, 96, 128, 136, 192, 256, 384])
        T = rng.choice([1, 2, 3, 5])
        H = rng.choice([1, 5, 16, 17, 33, 40])
        W = rng.choice([3, 8, 9, 16, 23, 52])
        kt = rng.choice([1, 3])
        cases.append((f"sweep{i}_c{cin}to{cout}_{T}x{H}x{W}_kt{kt}", cin, cout, (T, H, W), kt, rng.choice([0, 0, 1, 2])))
    return cases


@pytest.mark.parametrize("case", _sweep_cases(), ids=[c[0] for c in _sweep_cases()])
def test_conv3d_window_sweep(capi, case):
    name, cin, cout, (T, H, W), kt, impl = case
    g = torch.Generator(device="cpu").manual_seed(sum(name.encode()))
    x = _bf(torch.randn(cin, T, H, W, generator=g)).cuda()
    w = _bf(torch.randn(cout, cin, kt, 3, 3, generator=g) * (cin * kt * 9) ** -0.5).cuda()
    b = _bf(torch.randn(cout, generator=g) * 0.1).cuda()
    ref = _bf(F.conv3d(F.pad(x.float(), (1, 1, 1, 1, kt - 1, 0)).unsqueeze(0), w.float(), b.float())[0]).float()
    use_res = cout % 8 == 0
    res = _bf(torch.randn(cout, T, H, W, generator=g)).cuda() if use_res else None
    if use_res:
        ref = _bf(ref + res.float()).float()
    bias = torch.zeros((cout + 7) // 8 * 8, dtype=torch.bfloat16, device="cuda")
    bias[:cout] = b
    capi.conv_tuning(impl)            # 0: per shape (pair kernels), 1: tap form, 2: halo on single CTAs
    try:
        y, _ = capi.conv3d_cl(_cl(x), _w2d(w, cin), bias, kernel=(kt, 3, 3), pad=(kt - 1, 1, 1),
                              residual=_cl(res) if use_res else None, cout=cout)
        torch.cuda.synchronize()
    finally:
        capi.conv_tuning(0)
    got = y.float().permute(3, 0, 1, 2)
    assert float(got[cout:].abs().max()) == 0.0 if got.shape[0] > cout else True      # padded channels are exact zeros
    err = O.rel_l2(got[:cout], ref)
    assert err < 3e-3, f"{name} impl {impl}: rel_l2 {err:.3e}, max abs {float((got[:cout] - ref).abs().max()):.3e}"


def test_rowwise_kernels_against_torch(capi):
    g = torch.Generator(device="cpu").manual_seed(3)
    for C in (32, 96, 192, 384):
        x = _bf(torch.randn(5, 7, 11, C, generator=g) * 2).cuda()
        gamma = _bf(1.0 + 0.1 * torch.randn(C, generator=g)).cuda()
        for silu in (True, False):
            y = capi.vae_rmsnorm(x, gamma, silu=silu)
            want = F.normalize(x.float(), dim=-1) * C ** 0.5 * gamma.float()
            want = F.silu(want) if silu else want
            assert O.rel_l2(y.float(), want) < 4e-3, (C, silu)
    # nearest-exact 2x with and without the temporal interleave
    C, T, H, W = 32, 3, 5, 6
    x = _bf(torch.randn(T, H, W, C, generator=g)).cuda()
    up = capi.vae_upsample2x(x, None, T, H, W, C)
    want = F.interpolate(x.permute(0, 3, 1, 2).float(), scale_factor=(2.0, 2.0), mode="nearest-exact").permute(0, 2, 3, 1)
    assert torch.equal(up.float(), want)
    rest = _bf(torch.randn(T - 1, H, W, 2 * C, generator=g)).cuda()
    up3 = capi.vae_upsample2x(x, rest, 2 * T - 1, H, W, C)
    frames = [x[0]] + [rest[i // 2, :, :, (i % 2) * C:(i % 2 + 1) * C] for i in range(2 * (T - 1))]
    src = torch.stack(frames)
    want3 = F.interpolate(src.permute(0, 3, 1, 2).float(), scale_factor=(2.0, 2.0), mode="nearest-exact").permute(0, 2, 3, 1)
    assert torch.equal(up3.float(), want3)
    # softmax of fp32 scores with a padded tail
    L, Lp = 70, 96
    S = torch.randn(L, Lp, generator=g).cuda() * 4
    P = torch.empty(L, Lp, dtype=torch.bfloat16, device="cuda")
    capi.softmax_f32(S, P, L, Lp, 0.25)
    want = torch.softmax(S[:, :L] * 0.25, dim=-1)
    assert O.rel_l2(P[:, :L].float(), want) < 4e-3 and float(P[:, L:].abs().max()) == 0.0
    # layout changes and the latent affine
    z = _bf(torch.randn(16, 3, 4, 5, generator=g)).cuda()
    mean = torch.tensor(V.MEAN, device="cuda")
    inv_std = (1.0 / torch.tensor(V.STD)).cuda()
    cl = capi.vae_planes_to_cl(z, 16, mean=mean, inv_std=inv_std)
    want = _bf(_bf(z / _bf(inv_std).view(-1, 1, 1, 1)) + _bf(mean).view(-1, 1, 1, 1))
    assert torch.equal(cl.permute(3, 0, 1, 2), want)
    v3 = _bf(torch.randn(3, 2, 4, 6, generator=g)).cuda()
    cl3 = capi.vae_planes_to_cl(v3, 8)
    assert torch.equal(cl3[..., :3].permute(3, 0, 1, 2), v3) and float(cl3[..., 3:].abs().max()) == 0.0
    padded = torch.zeros(2, 4, 6 + 2, 8, dtype=torch.bfloat16, device="cuda")
    capi.vae_planes_to_cl(v3, 8, out=padded, wpad=1)
    assert torch.equal(padded[:, :, 1:-1, :3].permute(3, 0, 1, 2), v3)
    assert float(padded[:, :, 0].abs().max()) == 0.0 and float(padded[:, :, -1].abs().max()) == 0.0
    back = capi.vae_cl_to_planes(cl, 16, mean=mean, inv_std=inv_std)
    wantb = _bf(_bf(cl.permute(3, 0, 1, 2) - _bf(mean).view(-1, 1, 1, 1)) * _bf(inv_std).view(-1, 1, 1, 1))
    assert torch.equal(back, wantb)
    # head gather: nine taps x (3 + 1) partial-sum channels -> 3 planes
    T, H, W = 2, 5, 7
    Dp = _bf(torch.randn(T, H, W, 40, generator=g)).cuda()
    hb = torch.randn(3, generator=g).cuda()
    got = capi.vae_head_gather(Dp, hb, 3)
    want = torch.zeros(3, T, H, W, device="cuda")
    padded = F.pad(Dp.float().permute(3, 0, 1, 2), (1, 1, 1, 1))                 # (40, T, H+2, W+2)
    for dh in range(3):
        for dw in range(3):
            want += padded[(dh * 3 + dw) * 4:(dh * 3 + dw) * 4 + 3, :, dh:dh + H, dw:dw + W]
    want = _bf(want + hb.view(3, 1, 1, 1))
    assert float((got.float() - want.float()).abs().max()) <= 0.07 and O.rel_l2(got.float(), want.float()) < 3e-3   # <= 1 bf16 ulp
    # blending: two torch bf16 ops
    values = _bf(torch.randn(3, 2, 10, 12, generator=g)).cuda()
    tile = _bf(torch.randn(3, 2, 4, 5, generator=g)).cuda()
    mask = _bf(torch.rand(4, 5, generator=g)).cuda()
    want = values.clone()
    want[:, :, 3:7, 2:7] += tile * mask
    capi.vae_blend_(values, tile, mask, 3, 2)
    assert torch.equal(values, want)
    wgt = _bf(torch.rand(10, 12, generator=g) + 0.5).cuda()
    wantf = (want / wgt).clamp_(-1, 1)
    capi.vae_blend_finish_(values, wgt, clamp=True)
    assert torch.equal(values, wantf)


def _golden(golden_dir):
    g = torch.load(golden_dir / "vae.pt", weights_only=False)
    sd = V.random_state_dict(dim=g["dim"], seed=g["weight_seed"])
    return g, sd


def _bf16_oracle_err(fn, sd, x, want, **kw):
    """distance of the oracle's own bf16 run (on this GPU) to the fp32 result -- the tolerance budget"""
    sdb = {k: v.to(device="cuda", dtype=torch.bfloat16) for k, v in sd.items()}
    with torch.no_grad():
        return O.rel_l2(fn(sdb, x.to(device="cuda", dtype=torch.bfloat16), **kw).float().cpu(), want)


def test_decode_and_encode_against_reference_vectors(golden_dir):
    from goal_force_b200.wan_vae import WanVideoVAEB200
    g, sd = _golden(golden_dir)
    dim = g["dim"]
    vae = WanVideoVAEB200(sd, dim=dim)
    want = g["decode"].float()
    got = vae._decode_clip(g["z"][0].to(device="cuda", dtype=torch.bfloat16)).float().cpu().unsqueeze(0)
    err = O.rel_l2(got, want)
    budget = _bf16_oracle_err(V.decode, sd, g["z"], want, dim=dim)
    print(f"vae decode (dim {dim}): ours {err:.3e}  oracle-bf16 {budget:.3e}")
    assert err <= max(budget, 4e-3), (err, budget)
    want = g["encode"]
    got = vae._encode_clip(g["video"][0].to(device="cuda", dtype=torch.bfloat16)).float().cpu().unsqueeze(0)
    err = O.rel_l2(got, want)
    budget = _bf16_oracle_err(V.encode, sd, g["video"], want, dim=dim)
    print(f"vae encode (dim {dim}): ours {err:.3e}  oracle-bf16 {budget:.3e}")
    assert err <= max(budget, 4e-3), (err, budget)


def test_shipped_width_against_reference_vectors(golden_dir):
    """dim 96 against the vectors of the reference's own VideoVAE_ (tests/golden/vae.pt: decode96 / encode96)."""
    from goal_force_b200.wan_vae import WanVideoVAEB200
    g = torch.load(golden_dir / "vae.pt", weights_only=False)
    sd = V.random_state_dict(dim=96, seed=g["weight_seed96"])
    vae = WanVideoVAEB200(sd, dim=96)
    want = g["decode96"].float()
    got = vae._decode_clip(g["z96"][0].to(device="cuda", dtype=torch.bfloat16)).float().cpu().unsqueeze(0)
    err, budget = O.rel_l2(got, want), _bf16_oracle_err(V.decode, sd, g["z96"], want)
    print(f"vae decode96 vs reference vectors: ours {err:.3e}  oracle-bf16 {budget:.3e}")
    assert err <= max(budget, 4e-3), (err, budget)
    video = want.clamp(-1, 1)
    got_e = vae._encode_clip(video[0].to(device="cuda", dtype=torch.bfloat16)).float().cpu().unsqueeze(0)
    err_e, budget_e = O.rel_l2(got_e, g["encode96"]), _bf16_oracle_err(V.encode, sd, video, g["encode96"])
    print(f"vae encode96 vs reference vectors: ours {err_e:.3e}  oracle-bf16 {budget_e:.3e}")
    assert err_e <= max(budget_e, 4e-3), (err_e, budget_e)


def test_tiled_paths_against_reference_vectors(golden_dir):
    from goal_force_b200.wan_vae import WanVideoVAEB200
    g, sd = _golden(golden_dir)
    dim = g["dim"]
    vae = WanVideoVAEB200(sd, dim=dim)
    want = g["tiled_decode"].float()
    got = vae.decode(g["z_big"].to(torch.bfloat16), "cuda", tiled=True, tile_size=(4, 5), tile_stride=(3, 3)).float().cpu()
    err = O.rel_l2(got, want)
    budget = _bf16_oracle_err(V.tiled_decode, sd, g["z_big"], want, tile_size=(4, 5), tile_stride=(3, 3), dim=dim)
    print(f"vae tiled decode: ours {err:.3e}  oracle-bf16 {budget:.3e}")
    assert got.shape == want.shape and err <= max(budget, 5e-3), (err, budget)
    assert float(got.abs().max()) <= 1.0
    video = want.to(torch.bfloat16)
    wante = g["tiled_encode"]
    gote = vae.encode(video, "cuda", tiled=True, tile_size=(4, 5), tile_stride=(3, 3)).float().cpu()
    erre = O.rel_l2(gote, wante)
    budget = _bf16_oracle_err(V.tiled_encode, sd, want, wante, tile_size=(32, 40), tile_stride=(24, 24), dim=dim)
    print(f"vae tiled encode: ours {erre:.3e}  oracle-bf16 {budget:.3e}")
    assert gote.shape == wante.shape and erre <= max(budget, 5e-3), (erre, budget)
    # untiled public calls: decode clamps, encode of a batch of one
    one = vae.decode(g["z"].to(torch.bfloat16), "cuda").float().cpu()
    assert O.rel_l2(one, g["decode"].float().clamp(-1, 1)) < 2e-2


def test_real_width_clip_against_oracle():
    """dim 96 (the shipped Wan2.1 VAE width: 96/192/384 channels, two n-tiles, Cin % 64 != 0 k-blocks), a 9-frame
    64 x 96 clip: encode and decode vs the fp32 oracle on this GPU; budget = the oracle's own bf16 run."""
    from goal_force_b200.wan_vae import WanVideoVAEB200
    sd = V.random_state_dict(dim=96, seed=5)
    vae = WanVideoVAEB200(sd, dim=96)
    g = torch.Generator().manual_seed(11)
    z = torch.randn(1, 16, 3, 8, 12, generator=g)
    sdc = {k: v.cuda() for k, v in sd.items()}
    with torch.no_grad():
        want = V.decode(sdc, z.cuda()).cpu()
        got = vae._decode_clip(z[0].to(device="cuda", dtype=torch.bfloat16)).float().cpu().unsqueeze(0)
        err = O.rel_l2(got, want)
        budget = _bf16_oracle_err(V.decode, sd, z, want)
        print(f"vae decode (dim 96): ours {err:.3e}  oracle-bf16 {budget:.3e}")
        assert err <= max(budget, 4e-3), (err, budget)
        video = want.clamp(-1, 1)
        wante = V.encode(sdc, video.cuda()).cpu()
        gote = vae._encode_clip(video[0].to(device="cuda", dtype=torch.bfloat16)).float().cpu().unsqueeze(0)
        erre = O.rel_l2(gote, wante)
        budget = _bf16_oracle_err(V.encode, sd, video, wante)
        print(f"vae encode (dim 96): ours {erre:.3e}  oracle-bf16 {budget:.3e}")
        assert erre <= max(budget, 4e-3), (erre, budget)


def test_from_reference_module_is_a_drop_in(golden_dir):
    """INTEGRATION.md recipe: `pipe.vae = WanVideoVAEB200.from_reference(pipe.vae)` with an nn.Module that carries the
    reference's attribute tree, then the reference's call signatures (`encode(videos, device=, tiled=, tile_size=,
    tile_stride=)`, `decode(...)`) with the pipeline's argument meaning."""
    from goal_force_b200.wan_vae import WanVideoVAEB200
    from oracle.ref_standins import WanVideoVAEStandIn
    g, sd = _golden(golden_dir)
    ref_module = WanVideoVAEStandIn(sd).cuda()
    vae = WanVideoVAEB200.from_reference(ref_module, device="cuda")
    assert vae.dim == g["dim"] and vae.z_dim == 16 and vae.upsampling_factor == 8
    video = vae.decode(g["z_big"].to(torch.bfloat16), device="cuda", tiled=True, tile_size=(4, 5), tile_stride=(3, 3))
    assert video.shape == g["tiled_decode"].shape and video.is_cuda and video.dtype == torch.bfloat16
    assert O.rel_l2(video.float().cpu(), g["tiled_decode"].float()) < 2e-2
    lat = vae.encode([video[0]], device="cuda", tiled=True, tile_size=(4, 5), tile_stride=(3, 3))
    assert lat.shape == g["tiled_encode"].shape
    assert O.rel_l2(lat.float().cpu(), g["tiled_encode"]) < 2e-2
    # build_mask: the reference's formula (data tensor in, (1,1,1,H,W)-broadcastable mask out)
    m = vae.build_mask(torch.empty(1, 3, 5, 16, 24), (True, False, False, True), (8, 8))
    assert m.shape == (1, 1, 1, 16, 24) and torch.equal(m, V.build_mask(16, 24, (True, False, False, True), (8, 8)))


def test_job_driver_callables_on_the_b200_vae(golden_dir):
    """jobs.vae_control_encoder / jobs.vae_image_condition: the two VAE call sites that feed the denoiser
    (src/goal_force/wan_video_new.py:798-805 and :894-916) on WanVideoVAEB200, against the oracle."""
    from goal_force_b200 import jobs
    from goal_force_b200.pipeline import first_frame_mask
    from goal_force_b200.wan_vae import WanVideoVAEB200
    g, sd = _golden(golden_dir)
    dim = g["dim"]
    vae = WanVideoVAEB200(sd, dim=dim)
    gen = torch.Generator().manual_seed(9)
    control = torch.rand(9, 32, 48, 3, generator=gen).to(torch.bfloat16)            # (F, H, W, 3) as the dataset yields it
    lat = jobs.vae_control_encoder(vae, tiled=False)(control)
    with torch.no_grad():
        want = V.encode(sd, control.float().permute(3, 0, 1, 2).unsqueeze(0), dim=dim)
    assert lat.shape == (1, 16, 3, 4, 6) and lat.dtype == torch.bfloat16 and lat.is_cuda
    assert O.rel_l2(lat.float().cpu(), want) < 2e-2
    tiled = jobs.vae_control_encoder(vae, tiled=True, tile_size=(3, 4), tile_stride=(2, 2))(control)
    with torch.no_grad():
        want_t = V.tiled_encode(sd, control.float().permute(3, 0, 1, 2).unsqueeze(0), (24, 32), (16, 16), dim=dim)
    assert O.rel_l2(tiled.float().cpu(), want_t) < 2e-2
    image = (torch.rand(3, 32, 48, generator=gen) * 2 - 1).to(torch.bfloat16)
    y = jobs.vae_image_condition(vae, image, num_frames=9, tiled=False)
    assert y.shape == (1, 20, 3, 4, 6)
    assert torch.equal(y[0, :4].float().cpu(), first_frame_mask(9, 4, 6).float())
    clip = torch.zeros(1, 3, 9, 32, 48)
    clip[0, :, 0] = image.float()
    with torch.no_grad():
        want_y = V.encode(sd, clip, dim=dim)
    assert O.rel_l2(y[0, 4:].float().cpu(), want_y[0]) < 2e-2


def test_single_frame_and_odd_sizes(golden_dir):
    """T = 1 (an image: no temporal up / down-sampling happens, wan_video_vae.py:138-174), a 5-frame clip, and spatial
    sizes that are not multiples of the kernels' tiles -- against the fp32 oracle, budget = the oracle's own bf16 run."""
    from goal_force_b200.wan_vae import WanVideoVAEB200
    g, sd = _golden(golden_dir)
    dim = g["dim"]
    vae = WanVideoVAEB200(sd, dim=dim)
    gen = torch.Generator().manual_seed(21)
    for (T, h, w) in ((1, 5, 7), (2, 3, 11), (1, 9, 4)):
        z = torch.randn(1, 16, T, h, w, generator=gen)
        with torch.no_grad():
            want = V.decode(sd, z, dim=dim)
        got = vae.decode(z.to(torch.bfloat16), "cuda").float().cpu()
        assert got.shape == (1, 3, 4 * T - 3, 8 * h, 8 * w)
        err = O.rel_l2(got, want.clamp(-1, 1))
        budget = _bf16_oracle_err(lambda s_, x_, **kw: V.decode(s_, x_, **kw).clamp(-1, 1), sd, z, want.clamp(-1, 1), dim=dim)
        assert err <= max(budget, 5e-3), (T, h, w, err, budget)
        video = want.clamp(-1, 1)
        with torch.no_grad():
            want_e = V.encode(sd, video, dim=dim)
        got_e = vae.encode(video.to(torch.bfloat16), "cuda").float().cpu()
        assert got_e.shape == (1, 16, T, h, w)
        err_e = O.rel_l2(got_e, want_e)
        budget_e = _bf16_oracle_err(V.encode, sd, video, want_e, dim=dim)
        assert err_e <= max(budget_e, 5e-3), (T, h, w, err_e, budget_e)
