#!/usr/bin/env python
"""Kernel bring-up checks for a GPU box (run via gpurun). Each group runs in its own subprocess so a trapped kernel
(sticky CUDA error) cannot hide the remaining groups.

    python tools/gpu_check.py            # run every group, write gpurun_out/check_*.log + summary
    python tools/gpu_check.py gemm_small # run one group in-process
"""
from __future__ import annotations

import json
import math
import os
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
OUT = ROOT / "gpurun_out"


def _stats(got, ref):
    import torch
    g, r = got.float(), ref.float()
    diff = (g - r)
    rel = (diff.norm() / r.norm().clamp_min(1e-30)).item()
    return {"rel_l2": rel, "max_abs": diff.abs().max().item(), "ref_absmax": r.abs().max().item(),
            "nan": int(torch.isnan(g).sum().item())}


def _blockmap(got, ref, bs=32, tol=0.05):
    """coarse map of which (row-block, col-block) cells are wrong -- helps to read layout bugs"""
    import torch
    g, r = got.float(), ref.float()
    M, N = g.shape
    Mb, Nb = (M + bs - 1) // bs, (N + bs - 1) // bs
    lines = []
    for i in range(min(Mb, 16)):
        row = ""
        for j in range(min(Nb, 32)):
            gg = g[i * bs:(i + 1) * bs, j * bs:(j + 1) * bs]
            rr = r[i * bs:(i + 1) * bs, j * bs:(j + 1) * bs]
            e = (gg - rr).norm() / rr.norm().clamp_min(1e-6)
            row += "." if e < tol else ("N" if torch.isnan(gg).any() else "X")
        lines.append(row)
    return "\n".join(lines)


def timed(fn, iters=5, warmup=2):
    import torch
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    st, en = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    st.record()
    for _ in range(iters):
        fn()
    en.record()
    torch.cuda.synchronize()
    return st.elapsed_time(en) / iters


# ------------------------------------------------------------------------------------------------ groups
def g_gemm_small(cta_group=1):
    import torch
    from goal_force_b200 import capi
    torch.manual_seed(0)
    res = {}
    for (M, N, K) in [(128, 256, 64), (256, 256, 128), (300, 512, 320), (1000, 768, 144), (120, 64, 256)]:
        a = torch.randn(M, K, device="cuda").bfloat16()
        w = (torch.randn(N, K, device="cuda") / math.sqrt(K)).bfloat16()
        b = torch.randn(N, device="cuda").bfloat16()
        ref = a.float() @ w.float().t() + b.float()
        out = capi.gemm(a, w, b, cta_group=cta_group)
        torch.cuda.synchronize()
        st = _stats(out, ref)
        res[f"{M}x{N}x{K}"] = st
        print(f"gemm cg{cta_group} {M}x{N}x{K}: {st}", flush=True)
        if st["rel_l2"] > 1e-2:
            print(_blockmap(out, ref), flush=True)
    return res


def g_gemm_small2():
    return g_gemm_small(2)


def g_gemm_epi():
    import torch
    import torch.nn.functional as F
    from goal_force_b200 import capi
    torch.manual_seed(1)
    res = {}
    M, N, K = 520, 512, 256
    a = torch.randn(M, K, device="cuda").bfloat16()
    w = (torch.randn(N, K, device="cuda") / math.sqrt(K)).bfloat16()
    b = torch.randn(N, device="cuda").bfloat16()
    gate = torch.randn(N, device="cuda").bfloat16()
    x = torch.randn(M, N, device="cuda").bfloat16()
    lin = (a.float() @ w.float().t() + b.float())
    for cg in (1, 2):
        out = capi.gemm(a, w, b, epi=capi.GF_EPI_BIAS_GELU, cta_group=cg)
        res[f"gelu_cg{cg}"] = _stats(out, F.gelu(lin.bfloat16().float(), approximate="tanh"))
        out = capi.gemm(a, w, b, epi=capi.GF_EPI_BIAS_SILU, cta_group=cg)
        res[f"silu_cg{cg}"] = _stats(out, F.silu(lin.bfloat16().float()))
        out = capi.gemm(a, w, b, epi=capi.GF_EPI_GATE_RES, gate=gate, residual=x, cta_group=cg)
        ref = x.float() + (gate.float() * lin.bfloat16().float()).bfloat16().float()
        res[f"gate_cg{cg}"] = _stats(out, ref)
        xin = x.clone()
        capi.gemm(a, w, b, epi=capi.GF_EPI_GATE_RES, gate=None, residual=xin, out=xin, cta_group=cg)
        res[f"res_inplace_cg{cg}"] = _stats(xin, x.float() + lin.bfloat16().float())
    for k, v in res.items():
        print(k, v, flush=True)
    return res


def g_rowwise():
    import torch
    import torch.nn.functional as F
    from goal_force_b200 import capi
    torch.manual_seed(2)
    res = {}
    for d in (5120, 1536):
        rows = 777
        x = (torch.randn(rows, d, device="cuda") * 2 + 0.3).bfloat16()
        sh = torch.randn(d, device="cuda").bfloat16() * 0.5
        sc = torch.randn(d, device="cuda").bfloat16() * 0.5
        y = capi.layernorm(x, eps=1e-6, shift=sh, scale=sc)
        ref = (F.layer_norm(x.float(), (d,), eps=1e-6).bfloat16() * (1 + sc) + sh)
        res[f"ln_mod_{d}"] = _stats(y, ref)
        wt = torch.randn(d, device="cuda").bfloat16()
        bs = torch.randn(d, device="cuda").bfloat16()
        y = capi.layernorm(x, eps=1e-6, weight=wt, bias=bs)
        ref = F.layer_norm(x.float(), (d,), wt.float(), bs.float(), eps=1e-6)
        res[f"ln_affine_{d}"] = _stats(y, ref)
        # rmsnorm + rope
        heads = d // 128
        ang = torch.rand(rows, 64, device="cuda", dtype=torch.float64) * 6.28
        cs = torch.stack([ang.cos(), ang.sin()], dim=-1).float().contiguous()
        xx = x.clone()
        capi.rmsnorm_rope_(xx, wt, eps=1e-6, cos_sin=cs, head_dim=128)
        xf = x.float()
        n = (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-6)).bfloat16() * wt
        c = torch.view_as_complex(n.double().reshape(rows, heads, 64, 2))
        fr = torch.polar(torch.ones_like(ang), ang).unsqueeze(1)
        ref = torch.view_as_real(c * fr).flatten(1).bfloat16()
        res[f"rms_rope_{d}"] = _stats(xx, ref)
        res[f"rms_rope_{d}"]["mismatch_frac"] = (xx != ref).float().mean().item()
        xx = x.clone()
        capi.rmsnorm_rope_(xx, wt, eps=1e-6, cos_sin=None, head_dim=128)
        res[f"rms_{d}"] = _stats(xx, n)
        res[f"rms_{d}"]["mismatch_frac"] = (xx != n).float().mean().item()
    for k, v in res.items():
        print(k, v, flush=True)
    return res


def g_misc():
    import torch
    from goal_force_b200 import capi
    torch.manual_seed(3)
    res = {}
    F_, H, W = 3, 8, 12
    a = torch.randn(16, F_, H, W, device="cuda").bfloat16()
    b = torch.randn(20, F_, H, W, device="cuda").bfloat16()
    tok = capi.patch_gather(a, b)
    x = torch.cat([a, b], 0)  # (36,F,H,W)
    ref = x.reshape(36, F_, H // 2, 2, W // 2, 2).permute(1, 2, 4, 0, 3, 5).reshape(F_ * (H // 2) * (W // 2), 36 * 4)
    res["patch_gather_exact"] = bool(torch.equal(tok, ref))
    L = F_ * (H // 2) * (W // 2)
    t = torch.randn(L, 64, device="cuda").bfloat16()
    out = capi.unpatchify(t, 16, F_, H, W)
    ref = t.reshape(F_, H // 2, W // 2, 2, 2, 16).permute(5, 0, 1, 3, 2, 4).reshape(16, F_, H, W)
    res["unpatchify_exact"] = bool(torch.equal(out, ref))
    m = torch.randn(5, 6, 256, device="cuda").bfloat16()
    tm = torch.randn(6 * 256, device="cuda").bfloat16()
    res["add_rows_exact"] = bool(torch.equal(capi.add_rows(m.view(5, -1), tm), m.view(5, -1) + tm))
    v = torch.randn(5120, device="cuda").bfloat16()
    res["silu"] = _stats(capi.silu(v), torch.nn.functional.silu(v.float()))
    p, n, lat = (torch.randn(16, 5, 6, 8, device="cuda").bfloat16() for _ in range(3))
    got = capi.cfg_euler(p, n, lat, 5.0, -0.0371)
    pred = n + 5.0 * (p - n)
    ref = lat + pred * torch.tensor(-0.0371)
    res["cfg_euler_exact"] = bool(torch.equal(got, ref))
    res["cfg_euler"] = _stats(got, ref)
    xq = torch.randn(33, 8 * 128, device="cuda").bfloat16()
    pk = capi.ulysses_pack(xq, 8, 128, 4)
    ref = xq.view(33, 4, 2, 128).permute(1, 0, 2, 3).reshape(4, 33, 256)
    res["ulysses_pack_exact"] = bool(torch.equal(pk, ref))
    res["ulysses_unpack_exact"] = bool(torch.equal(capi.ulysses_unpack(pk, 33, 8, 128, 4), xq))
    for k, v in res.items():
        print(k, v, flush=True)
    return res


def _attn_ref(q, k, v, heads):
    import torch
    Lq, Lk = q.shape[0], k.shape[0]
    qh = q.float().view(Lq, heads, 128).transpose(0, 1)
    kh = k.float().view(Lk, heads, 128).transpose(0, 1)
    vh = v.float().view(Lk, heads, 128).transpose(0, 1)
    o = torch.nn.functional.scaled_dot_product_attention(qh[None], kh[None], vh[None])[0]
    return o.transpose(0, 1).reshape(Lq, heads * 128)


def g_attn_small():
    import torch
    from goal_force_b200 import capi
    torch.manual_seed(4)
    res = {}
    for (Lq, Lk, heads, amp) in [(256, 128, 1, 1.0), (256, 256, 2, 1.0), (300, 200, 2, 1.0), (512, 1024, 3, 1.0),
                                 (256, 512, 1, 4.0), (1000, 1333, 2, 2.0)]:
        q = (torch.randn(Lq, heads * 128, device="cuda") * amp).bfloat16()
        k = (torch.randn(Lk, heads * 128, device="cuda") * amp).bfloat16()
        v = torch.randn(Lk, heads * 128, device="cuda").bfloat16()
        o = capi.attention(q, k, v, heads)
        torch.cuda.synchronize()
        ref = _attn_ref(q, k, v, heads)
        st = _stats(o, ref)
        res[f"Lq{Lq}_Lk{Lk}_h{heads}_a{amp}"] = st
        print(f"attn Lq={Lq} Lk={Lk} h={heads} amp={amp}: {st}", flush=True)
        if st["rel_l2"] > 2e-2:
            print(_blockmap(o, ref, tol=0.1), flush=True)
    # fused-qkv strided views
    L, heads = 384, 2
    qkv = torch.randn(L, 3 * heads * 128, device="cuda").bfloat16()
    d = heads * 128
    o = capi.attention(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], heads)
    res["strided_qkv"] = _stats(o, _attn_ref(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], heads))
    print("strided", res["strided_qkv"], flush=True)
    return res


def g_gemm_perf():
    import torch
    from goal_force_b200 import capi
    res = {}
    M = 32760
    for (N, K) in [(5120, 5120), (15360, 5120), (13824, 5120), (5120, 13824)]:
        a = torch.randn(M, K, device="cuda").bfloat16()
        w = (torch.randn(N, K, device="cuda") / math.sqrt(K)).bfloat16()
        b = torch.randn(N, device="cuda").bfloat16()
        out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        fl = 2.0 * M * N * K
        for cg in (1, 2):
            ms = timed(lambda: capi.gemm(a, w, b, out=out, cta_group=cg))
            res[f"N{N}_K{K}_cg{cg}"] = {"ms": ms, "tflops": fl / ms / 1e9}
            print(f"gemm {M}x{N}x{K} cg{cg}: {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s", flush=True)
        ms = timed(lambda: torch.nn.functional.linear(a, w, b))
        res[f"N{N}_K{K}_cublas"] = {"ms": ms, "tflops": fl / ms / 1e9}
        print(f"cublas {M}x{N}x{K}: {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s", flush=True)
        # spot-check correctness at full size on a row sample
        capi.gemm(a, w, b, out=out, cta_group=2)
        idx = torch.randint(0, M, (64,), device="cuda")
        ref = a[idx].float() @ w.float().t() + b.float()
        res[f"N{N}_K{K}_check"] = _stats(out[idx], ref)
        print("  check", res[f"N{N}_K{K}_check"], flush=True)
        del a, w, out
    return res


def g_gemm_sustained():
    """sustained (power-capped) GEMM rate at the four block shapes, for the rasterisation set by GF_GEMM_GROUP_M"""
    import torch
    from goal_force_b200 import capi
    res = {"group_m": os.environ.get("GF_GEMM_GROUP_M", "default")}
    M = 32760
    tot_ms = 0.0
    for (N, K) in [(15360, 5120), (5120, 5120), (13824, 5120), (5120, 13824)]:
        a = torch.randn(M, K, device="cuda").bfloat16()
        w = (torch.randn(N, K, device="cuda") / math.sqrt(K)).bfloat16()
        b = torch.randn(N, device="cuda").bfloat16()
        out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        ms = timed(lambda: capi.gemm(a, w, b, out=out), iters=150, warmup=10)
        res[f"N{N}_K{K}"] = {"ms": round(ms, 4), "tflops": round(2.0 * M * N * K / ms / 1e9, 1)}
        tot_ms += ms
        del a, w, out
    res["sum_ms"] = round(tot_ms, 4)
    print(res, flush=True)
    return res


def g_attn_perf():
    import torch
    from goal_force_b200 import capi
    res = {}
    heads = 40
    for L in (8192, 32760):
        qkv = torch.randn(L, 3 * heads * 128, device="cuda").bfloat16()
        d = heads * 128
        o = torch.empty(L, d, device="cuda", dtype=torch.bfloat16)
        fl = 4.0 * L * L * d
        ms = timed(lambda: capi.attention(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], heads, out=o), iters=3, warmup=1)
        res[f"self_L{L}"] = {"ms": ms, "tflops": fl / ms / 1e9}
        print(f"attn self L={L}: {ms:.3f} ms {fl / ms / 1e9:.1f} TFLOP/s", flush=True)
        try:
            from flash_attn import flash_attn_func
            q4 = qkv[:, :d].reshape(1, L, heads, 128)
            k4 = qkv[:, d:2 * d].reshape(1, L, heads, 128)
            v4 = qkv[:, 2 * d:].reshape(1, L, heads, 128)
            ms2 = timed(lambda: flash_attn_func(q4, k4, v4), iters=3, warmup=1)
            res[f"fa2_L{L}"] = {"ms": ms2, "tflops": fl / ms2 / 1e9}
            print(f"flash_attn2 L={L}: {ms2:.3f} ms {fl / ms2 / 1e9:.1f} TFLOP/s", flush=True)
            ref = flash_attn_func(q4, k4, v4).reshape(L, d)
            res[f"vs_fa2_L{L}"] = _stats(o, ref)
            print("  vs fa2", res[f"vs_fa2_L{L}"], flush=True)
        except Exception as e:  # noqa: BLE001
            print("flash_attn unavailable:", repr(e)[:200], flush=True)
        if L == 8192:
            ref = _attn_ref(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], heads)
            res[f"check_L{L}"] = _stats(o, ref)
            print("  check", res[f"check_L{L}"], flush=True)
    # cross attention shape
    L = 32760
    q = torch.randn(L, heads * 128, device="cuda").bfloat16()
    kv = torch.randn(512, 2 * heads * 128, device="cuda").bfloat16()
    o = torch.empty(L, heads * 128, device="cuda", dtype=torch.bfloat16)
    ms = timed(lambda: capi.attention(q, kv[:, :heads * 128], kv[:, heads * 128:], heads, out=o))
    fl = 4.0 * L * 512 * heads * 128
    res["cross"] = {"ms": ms, "tflops": fl / ms / 1e9}
    print(f"attn cross: {ms:.3f} ms {fl / ms / 1e9:.1f} TFLOP/s", flush=True)
    return res


def g_rowwise_perf():
    import torch
    from goal_force_b200 import capi
    res = {}
    L, d = 32760, 5120
    x = torch.randn(L, d, device="cuda").bfloat16()
    y = torch.empty_like(x)
    sh = torch.randn(d, device="cuda").bfloat16()
    sc = torch.randn(d, device="cuda").bfloat16()
    ms = timed(lambda: capi.layernorm(x, eps=1e-6, shift=sh, scale=sc, out=y), iters=20)
    res["ln_mod"] = {"ms": ms, "gbs": 2 * L * d * 2 / ms / 1e6}
    cs = torch.randn(L, 64, 2, device="cuda")
    ms = timed(lambda: capi.rmsnorm_rope_(x, sh, eps=1e-6, cos_sin=cs, head_dim=128), iters=20)
    res["rms_rope"] = {"ms": ms, "gbs": 2 * L * d * 2 / ms / 1e6}
    for k, v in res.items():
        print(k, v, flush=True)
    return res


def g_attn_one():
    """accuracy + sustained timing of the attention kernel as configured by GF_ATTN_EMU_PAIRS (one process = one variant)"""
    import torch
    from goal_force_b200 import capi
    res = {"emu_pairs": os.environ.get("GF_ATTN_EMU_PAIRS", "default"), "impl": os.environ.get("GF_ATTN_IMPL", "default")}
    heads, d = 40, 5120
    torch.manual_seed(0)
    # accuracy: L = 4096, all heads, vs fp32 SDPA; plus a growing-magnitude case that forces the rescale branch
    L = 4096
    qkv = torch.randn(L, 3 * d, device="cuda").bfloat16()
    o = capi.attention(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], heads)
    res["acc_L4096"] = _stats(o, _attn_ref(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], heads))
    q2 = torch.randn(1000, 2 * 128, device="cuda")
    k2 = torch.randn(1333, 2 * 128, device="cuda") * torch.linspace(0.2, 6.0, 1333, device="cuda")[:, None]
    v2 = torch.randn(1333, 2 * 128, device="cuda")
    q2, k2, v2 = q2.bfloat16(), k2.bfloat16(), v2.bfloat16()
    res["acc_rescale"] = _stats(capi.attention(q2, k2, v2, 2), _attn_ref(q2, k2, v2, 2))
    q3 = (torch.randn(300, 128, device="cuda") * 3).bfloat16()
    k3 = (torch.randn(200, 128, device="cuda") * 3).bfloat16()
    v3 = torch.randn(200, 128, device="cuda").bfloat16()
    res["acc_ragged_peaky"] = _stats(capi.attention(q3, k3, v3, 1), _attn_ref(q3, k3, v3, 1))
    L = 32760
    qkv = torch.randn(L, 3 * d, device="cuda").bfloat16()
    o = torch.empty(L, d, device="cuda", dtype=torch.bfloat16)
    fl = 4.0 * L * L * d
    run = lambda: capi.attention(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], heads, out=o)  # noqa: E731
    ms = timed(run, iters=3, warmup=1)
    res["burst"] = {"ms": ms, "tflops": fl / ms / 1e9}
    ms = timed(run, iters=40, warmup=0)          # ~0.7 s back to back: settles under the power cap
    res["sustained"] = {"ms": ms, "tflops": fl / ms / 1e9}
    # context: torch's own fused attention on this GPU (F.scaled_dot_product_attention; backend chosen by torch)
    try:
        import torch.nn.functional as F
        q4, k4, v4 = (qkv[:, i * d:(i + 1) * d].reshape(1, L, heads, 128).transpose(1, 2) for i in range(3))
        sd = lambda: F.scaled_dot_product_attention(q4, k4, v4)  # noqa: E731
        ms = timed(sd, iters=3, warmup=1)
        res["torch_sdpa_burst"] = {"ms": ms, "tflops": fl / ms / 1e9}
        ms = timed(sd, iters=40, warmup=0)
        res["torch_sdpa_sustained"] = {"ms": ms, "tflops": fl / ms / 1e9}
        from torch.nn.attention import SDPBackend, sdpa_kernel
        for name, be in (("cudnn", SDPBackend.CUDNN_ATTENTION), ("flash", SDPBackend.FLASH_ATTENTION)):
            try:
                with sdpa_kernel(be):
                    ms = timed(sd, iters=3, warmup=1)
                res[f"torch_sdpa_{name}"] = {"ms": ms, "tflops": fl / ms / 1e9}
            except Exception as e:  # noqa: BLE001
                res[f"torch_sdpa_{name}"] = {"error": repr(e)[:120]}
    except Exception as e:  # noqa: BLE001
        res["torch_sdpa_burst"] = {"error": repr(e)[:200]}
    idx = torch.randint(0, L, (48,), device="cuda")
    ref = _attn_ref(qkv[idx, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], heads)
    res["acc_L32760_rows"] = _stats(o[idx], ref)
    # the 8-GPU Ulysses shape: 5 heads per rank, full sequence (640 work items on 148 SMs -> tail splitting)
    h8 = 5
    ms = timed(lambda: capi.attention(qkv[:, :h8 * 128], qkv[:, d:d + h8 * 128], qkv[:, 2 * d:2 * d + h8 * 128], h8,
                                      out=o[:, :h8 * 128]), iters=10, warmup=2)
    res["sp8_shape"] = {"ms": ms, "tflops": 4.0 * L * L * h8 * 128 / ms / 1e9}
    kv = torch.randn(512, 2 * d, device="cuda").bfloat16()
    q = qkv[:, :d].contiguous()
    ms = timed(lambda: capi.attention(q, kv[:, :d], kv[:, d:], heads, out=o), iters=20)
    res["cross"] = {"ms": ms, "tflops": 4.0 * L * 512 * d / ms / 1e9}
    for k, v in res.items():
        print(k, v, flush=True)
    return res


def g_attn_sweep():
    res = {}
    for var in os.environ.get("GF_ATTN_SWEEP", "80:0,128:4").split(","):
        impl, emu = var.split(":")
        env = dict(os.environ, GF_ATTN_IMPL=impl, GF_ATTN_EMU_PAIRS=emu, PYTHONUNBUFFERED="1")
        r = subprocess.run([sys.executable, __file__, "attn_one"], env=env, capture_output=True, text=True, timeout=280)
        print(f"---- GF_ATTN_IMPL={impl} GF_ATTN_EMU_PAIRS={emu} rc={r.returncode}\n{r.stdout[-2500:]}{r.stderr[-1500:]}", flush=True)
        try:
            res[var] = json.loads((OUT / "check_attn_one.json").read_text())
        except Exception:  # noqa: BLE001
            res[var] = {"rc": r.returncode}
    return res


def g_attn_ab():
    """Same-process, same-box A/B of attention kernel variants at the config-2 self-attention shape
    (GF_ATTN_AB = "impl:emu,impl:emu,..."): accuracy on the rescale / ragged cases, then burst and sustained timings,
    visiting the variants round-robin twice so that thermal drift hits all of them alike; cuDNN SDPA beside it."""
    import torch
    import torch.nn.functional as F
    from goal_force_b200 import capi
    variants = [tuple(int(x) for x in v.split(":")) for v in
                os.environ.get("GF_ATTN_AB", "80:0,81:0,82:0,80:2,81:2,82:2,82:4").split(",")]
    heads, d, L = 40, 5120, 32760
    torch.manual_seed(0)
    res = {}
    q2 = torch.randn(1000, 2 * 128, device="cuda").bfloat16()
    k2 = (torch.randn(1333, 2 * 128, device="cuda") * torch.linspace(0.2, 6.0, 1333, device="cuda")[:, None]).bfloat16()
    v2 = torch.randn(1333, 2 * 128, device="cuda").bfloat16()
    q3 = (torch.randn(300, 128, device="cuda") * 3).bfloat16()
    k3 = (torch.randn(200, 128, device="cuda") * 3).bfloat16()
    v3 = torch.randn(200, 128, device="cuda").bfloat16()
    qa = torch.randn(4096, 3 * d, device="cuda").bfloat16()
    refs = (_attn_ref(q2, k2, v2, 2), _attn_ref(q3, k3, v3, 1), _attn_ref(qa[:, :d], qa[:, d:2 * d], qa[:, 2 * d:], heads))
    for impl, emu in variants:
        capi.attention_tuning(impl, emu)
        key = f"{impl}:{emu}"
        res[key] = {"acc_rescale": _stats(capi.attention(q2, k2, v2, 2), refs[0])["rel_l2"],
                    "acc_ragged": _stats(capi.attention(q3, k3, v3, 1), refs[1])["rel_l2"],
                    "acc_L4096": _stats(capi.attention(qa[:, :d], qa[:, d:2 * d], qa[:, 2 * d:], heads), refs[2])["rel_l2"],
                    "burst": [], "sustained": []}
        print(key, res[key], flush=True)
    qkv = torch.randn(L, 3 * d, device="cuda").bfloat16()
    o = torch.empty(L, d, device="cuda", dtype=torch.bfloat16)
    fl = 4.0 * L * L * d
    run = lambda: capi.attention(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], heads, out=o)  # noqa: E731
    q4, k4, v4 = (qkv[:, i * d:(i + 1) * d].reshape(1, L, heads, 128).transpose(1, 2) for i in range(3))
    sd = lambda: F.scaled_dot_product_attention(q4, k4, v4)  # noqa: E731
    res["cudnn_sdpa"] = {"burst": [], "sustained": []}
    for rnd in range(2):
        for impl, emu in variants:
            capi.attention_tuning(impl, emu)
            key = f"{impl}:{emu}"
            torch.cuda.synchronize(); time.sleep(0.5)
            res[key]["burst"].append(round(fl / timed(run, iters=3, warmup=1) / 1e9, 1))
            res[key]["sustained"].append(round(fl / timed(run, iters=60, warmup=0) / 1e9, 1))
            print(key, "burst", res[key]["burst"], "sustained", res[key]["sustained"], flush=True)
        torch.cuda.synchronize(); time.sleep(0.5)
        res["cudnn_sdpa"]["burst"].append(round(fl / timed(sd, iters=3, warmup=1) / 1e9, 1))
        res["cudnn_sdpa"]["sustained"].append(round(fl / timed(sd, iters=60, warmup=0) / 1e9, 1))
        print("cudnn_sdpa", res["cudnn_sdpa"], flush=True)
    capi.attention_tuning(0, -1)
    return res


GROUPS = {
    "gemm_small": g_gemm_small, "gemm_small2": g_gemm_small2, "gemm_epi": g_gemm_epi, "rowwise": g_rowwise,
    "misc": g_misc, "attn_small": g_attn_small, "gemm_perf": g_gemm_perf, "attn_perf": g_attn_perf,
    "rowwise_perf": g_rowwise_perf, "attn_one": g_attn_one, "attn_sweep": g_attn_sweep,
    "gemm_sustained": g_gemm_sustained, "attn_ab": g_attn_ab,
}


def main():
    OUT.mkdir(exist_ok=True)
    if len(sys.argv) > 1 and sys.argv[1] != "--only":
        name = sys.argv[1]
        res = GROUPS[name]()
        (OUT / f"check_{name}.json").write_text(json.dumps(res, indent=1))
        return
    names = sys.argv[2].split(",") if len(sys.argv) > 2 else list(GROUPS)
    summary = {}
    for name in names:
        t0 = time.time()
        log = OUT / f"check_{name}.log"
        with open(log, "w") as f:
            try:
                r = subprocess.run([sys.executable, __file__, name], stdout=f, stderr=subprocess.STDOUT, timeout=1500,
                                   env=dict(os.environ, PYTHONUNBUFFERED="1"))
                rc = r.returncode
            except subprocess.TimeoutExpired:
                rc = "timeout"
        summary[name] = {"rc": rc, "sec": round(time.time() - t0, 1)}
        print(f"[{name}] rc={rc} {summary[name]['sec']}s", flush=True)
        print(log.read_text()[-3000:], flush=True)
    (OUT / "check_summary.json").write_text(json.dumps(summary, indent=1))


if __name__ == "__main__":
    main()
