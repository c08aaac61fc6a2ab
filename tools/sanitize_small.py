#!/usr/bin/env python
"""Every C-ABI entry point once at small shapes -- the workload for compute-sanitizer (memcheck / racecheck):

    compute-sanitizer --tool memcheck  python tools/sanitize_small.py
    compute-sanitizer --tool racecheck python tools/sanitize_small.py

Shapes are ragged on purpose (rows not multiples of the tiles, a partial last kv block) so that every bounds check of
the kernels is exercised; results are compared with torch so a wrong answer also fails the run.
"""
import math
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from goal_force_b200 import capi  # noqa: E402


def rel(a, b):
    return float((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-30))


def main():
    torch.manual_seed(0)
    dev = "cuda"
    ok = True

    def check(name, err, tol):
        nonlocal ok
        good = err <= tol
        ok &= good
        print(f"{name:28s} rel {err:.3e} {'ok' if good else 'FAIL'}", flush=True)

    only_vae = "--vae-only" in sys.argv
    if not only_vae:
        M, N, K = 300, 512, 320
        a = torch.randn(M, K, device=dev).bfloat16()
        w = (torch.randn(N, K, device=dev) / math.sqrt(K)).bfloat16()
        b = torch.randn(N, device=dev).bfloat16()
        x = torch.randn(M, N, device=dev).bfloat16()
        g = torch.randn(N, device=dev).bfloat16()
        lin = a.float() @ w.float().t() + b.float()
        for cg in (1, 2):
            check(f"gemm bias cg{cg}", rel(capi.gemm(a, w, b, cta_group=cg), lin), 1e-2)
            check(f"gemm gelu cg{cg}", rel(capi.gemm(a, w, b, epi=capi.GF_EPI_BIAS_GELU, cta_group=cg),
                                           F.gelu(lin.bfloat16().float(), approximate="tanh")), 1e-2)
            check(f"gemm gate_res cg{cg}", rel(capi.gemm(a, w, b, epi=capi.GF_EPI_GATE_RES, gate=g, residual=x, cta_group=cg),
                                               x.float() + (g.float() * lin.bfloat16().float()).bfloat16().float()), 1e-2)
        d = 1536
        xr = (torch.randn(77, d, device=dev) * 2 + 0.3).bfloat16()
        sh, sc = torch.randn(d, device=dev).bfloat16() * 0.5, torch.randn(d, device=dev).bfloat16() * 0.5
        check("layernorm modulate", rel(capi.layernorm(xr, eps=1e-6, shift=sh, scale=sc),
                                        F.layer_norm(xr.float(), (d,), eps=1e-6).bfloat16() * (1 + sc) + sh), 1e-2)
        wt = torch.randn(d, device=dev).bfloat16()
        cs = torch.randn(77, 64, 2, device=dev)
        xx = xr.clone()
        capi.rmsnorm_rope_(xx, wt, eps=1e-6, cos_sin=cs, head_dim=128)
        qkv = torch.randn(77, 3 * d, device=dev).bfloat16()
        capi.qk_rmsnorm_rope_(qkv, wt, wt, eps=1e-6, cos_sin=cs, head_dim=128)
        print("rmsnorm_rope / qk_rmsnorm_rope ran", flush=True)
        for impl in (80, 128):
            capi.attention_tuning(impl, -1)
            for (Lq, Lk, h) in ((300, 200, 2), (130, 81, 1), (257, 512, 1)):
                q = torch.randn(Lq, h * 128, device=dev).bfloat16()
                k = torch.randn(Lk, h * 128, device=dev).bfloat16()
                v = torch.randn(Lk, h * 128, device=dev).bfloat16()
                o = capi.attention(q, k, v, h)
                qh, kh, vh = (t.float().view(t.shape[0], h, 128).transpose(0, 1)[None] for t in (q, k, v))
                ref = F.scaled_dot_product_attention(qh, kh, vh)[0].transpose(0, 1).reshape(Lq, h * 128)
                check(f"attention impl{impl} {Lq}x{Lk}x{h}", rel(o, ref), 5e-3)
        capi.attention_tuning(0, -1)
        s0 = torch.randn(16, 3, 8, 12, device=dev).bfloat16()
        s1 = torch.randn(20, 3, 8, 12, device=dev).bfloat16()
        tok = capi.patch_gather(s0, s1)
        ref = torch.cat([s0, s1], 0).reshape(36, 3, 4, 2, 6, 2).permute(1, 2, 4, 0, 3, 5).reshape(72, 144)
        check("patch_gather", 0.0 if torch.equal(tok, ref) else 1.0, 0.0)
        t = torch.randn(72, 64, device=dev).bfloat16()
        out = capi.unpatchify(t, 16, 3, 8, 12)
        check("unpatchify", 0.0 if torch.equal(out, t.reshape(3, 4, 6, 2, 2, 16).permute(5, 0, 1, 3, 2, 4).reshape(16, 3, 8, 12)) else 1.0, 0.0)
        m = torch.randn(5, 6 * 256, device=dev).bfloat16()
        tm = torch.randn(6 * 256, device=dev).bfloat16()
        check("add_rows", 0.0 if torch.equal(capi.add_rows(m, tm), m + tm) else 1.0, 0.0)
        y = torch.randn(1003, device=dev).bfloat16()[:1000].contiguous()
        z = torch.randn(1000, device=dev).bfloat16()
        want = y + z
        check("add_", 0.0 if torch.equal(capi.add_(y, z), want) else 1.0, 0.0)
        check("silu", rel(capi.silu(z), F.silu(z.float())), 1e-2)
        p, n, lat = (torch.randn(16, 5, 6, 8, device=dev).bfloat16() for _ in range(3))
        want = lat + (n + 5.0 * (p - n)) * torch.tensor(-0.0371)
        check("cfg_euler", 0.0 if torch.equal(capi.cfg_euler(p, n, lat, 5.0, -0.0371), want) else 1.0, 0.0)
        capi.timestep_embedding(torch.tensor([937.0], device=dev).bfloat16(), 256)
        xq = torch.randn(33, 8 * 128, device=dev).bfloat16()
        pk = capi.ulysses_pack(xq, 8, 128, 4)
        check("ulysses pack/unpack", 0.0 if torch.equal(capi.ulysses_unpack(pk, 33, 8, 128, 4), xq) else 1.0, 0.0)
    # ---- Wan VAE pieces: both convolution kernels (ragged frames), strided maps, streaming kernels, a tiny clip
    torch.backends.cudnn.allow_tf32 = False

    def conv_case(name, cin, cout, dims, kernel, stride, pad, fused):
        T, H, W = dims
        taps = kernel[0] * kernel[1] * kernel[2]
        xc = torch.randn(cin, T, H, W, device=dev).bfloat16()
        wc = (torch.randn(cout, cin, *kernel, device=dev) / math.sqrt(cin * taps)).bfloat16()
        bc = torch.zeros((cout + 7) // 8 * 8, device=dev).bfloat16()
        if stride[1] == 2:
            xp, Ho, Wo = F.pad(xc.float(), (0, 1, 0, 1, pad[0], 0)), H // 2, W // 2
        else:
            xp = F.pad(xc.float(), (pad[2], kernel[2] - 1 - pad[2], pad[1], kernel[1] - 1 - pad[1], pad[0], 0))
            Ho, Wo = H, W
        want = F.conv3d(xp[None], wc.float(), None, stride=stride)[0]
        gam = torch.ones((cout + 7) // 8 * 8, device=dev).bfloat16() if fused else None
        res = torch.randn(want.shape[1], Ho, Wo, cout, device=dev).bfloat16() if fused else None
        yv, _ = capi.conv3d_cl(xc.permute(1, 2, 3, 0).contiguous(), wc.permute(0, 2, 3, 4, 1).reshape(cout, -1).contiguous(),
                               bc, kernel=kernel, stride=stride, pad=pad, out_dims=(want.shape[1], Ho, Wo), residual=res,
                               gamma=gam, cout=cout)
        if fused:
            want = want.bfloat16().float() + res.permute(3, 0, 1, 2).float()
        check(name, rel(yv.permute(3, 0, 1, 2)[:cout], want), 6e-3)

    conv_case("conv halo 96>96 fused", 96, 96, (3, 19, 13), (3, 3, 3), (1, 1, 1), (2, 1, 1), True)
    conv_case("conv halo 16>32", 16, 32, (2, 9, 11), (3, 3, 3), (1, 1, 1), (2, 1, 1), False)
    conv_case("conv taps 192>192 fused", 192, 192, (3, 9, 13), (3, 3, 3), (1, 1, 1), (2, 1, 1), True)
    conv_case("conv taps 192>384", 192, 384, (2, 7, 9), (3, 3, 3), (1, 1, 1), (2, 1, 1), False)
    conv_case("conv stride2 96>96", 96, 96, (2, 10, 14), (1, 3, 3), (1, 2, 2), (0, 0, 0), False)
    conv_case("conv time stride2", 64, 64, (7, 5, 6), (3, 1, 1), (2, 1, 1), (0, 0, 0), False)
    from goal_force_b200.wan_vae import WanVideoVAEB200
    from oracle import wan_vae_oracle as V
    sd = V.random_state_dict(dim=32, seed=0)
    vae = WanVideoVAEB200(sd, dim=32)
    zz = torch.randn(1, 16, 2, 5, 6)
    with torch.no_grad():
        check("vae decode (dim 32)", rel(vae._decode_clip(zz[0].to(dev).bfloat16()).cpu(), V.decode(sd, zz, dim=32)[0]), 3e-2)
        vid = torch.randn(1, 3, 5, 24, 40).clamp(-1, 1)
        check("vae encode (dim 32)", rel(vae._encode_clip(vid[0].to(dev).bfloat16()).cpu(), V.encode(sd, vid, dim=32)[0]), 3e-2)
        check("vae tiled decode", rel(vae.decode(zz.bfloat16(), dev, tiled=True, tile_size=(4, 4), tile_stride=(2, 3)).cpu(),
                                      V.tiled_decode(sd, zz, (4, 4), (2, 3), dim=32)), 3e-2)
    torch.cuda.synchronize()
    print("ALL OK" if ok else "FAILURES", flush=True)
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
