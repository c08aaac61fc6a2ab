"""Per-kernel parity on the GPU, through the C ABI (goal_force_b200.capi -> libgoalforce_b200.so).
Floating-point kernels are compared against a plain fp32 torch restatement of the same op; index / rounding-chain
kernels (patch gather, unpatchify, modulation add, CFG+Euler, Ulysses pack) must be bit-exact."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

# bf16 output rounding alone gives ~1.7e-3 relative L2 against an fp32 reference
BF16_ROUND_TOL = 3e-3


def rel(a, b):
    a, b = a.float(), b.float()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


@pytest.mark.parametrize("cta_group", [1, 2])
@pytest.mark.parametrize("shape", [(128, 256, 64), (1, 5120, 256), (300, 512, 320), (1000, 768, 144), (120, 64, 256),
                                   (2050, 1536, 1536)])
def test_gemm_bias(capi, cta_group, shape):
    M, N, K = shape
    torch.manual_seed(0)
    a = torch.randn(M, K, device="cuda").bfloat16()
    w = (torch.randn(N, K, device="cuda") / math.sqrt(K)).bfloat16()
    b = torch.randn(N, device="cuda").bfloat16()
    out = capi.gemm(a, w, b, cta_group=cta_group)
    assert rel(out, a.float() @ w.float().t() + b.float()) < BF16_ROUND_TOL


@pytest.mark.parametrize("cta_group", [1, 2])
def test_gemm_epilogues(capi, cta_group):
    torch.manual_seed(1)
    M, N, K = 520, 512, 256
    a = torch.randn(M, K, device="cuda").bfloat16()
    w = (torch.randn(N, K, device="cuda") / math.sqrt(K)).bfloat16()
    b = torch.randn(N, device="cuda").bfloat16()
    gate = torch.randn(N, device="cuda").bfloat16()
    x = torch.randn(M, N, device="cuda").bfloat16()
    lin = (a.float() @ w.float().t() + b.float()).bfloat16().float()
    assert rel(capi.gemm(a, w, b, epi=capi.GF_EPI_BIAS_GELU, cta_group=cta_group),
               F.gelu(lin, approximate="tanh")) < BF16_ROUND_TOL
    assert rel(capi.gemm(a, w, b, epi=capi.GF_EPI_BIAS_SILU, cta_group=cta_group), F.silu(lin)) < BF16_ROUND_TOL
    ref = x.float() + (gate.float() * lin).bfloat16().float()
    assert rel(capi.gemm(a, w, b, epi=capi.GF_EPI_GATE_RES, gate=gate, residual=x, cta_group=cta_group),
               ref) < BF16_ROUND_TOL
    xin = x.clone()   # in place, no gate (cross-attention residual / zero-conv inject)
    capi.gemm(a, w, b, epi=capi.GF_EPI_GATE_RES, residual=xin, out=xin, cta_group=cta_group)
    assert rel(xin, x.float() + lin) < BF16_ROUND_TOL


def test_gemm_linearity_at_full_width(capi):
    # size-independent property at the A14B shapes: gemm(a1 + a2) == gemm(a1) + gemm(a2) (no bias), sampled rows
    torch.manual_seed(2)
    M, N, K = 4096, 5120, 5120
    a1 = torch.randn(M, K, device="cuda").bfloat16()
    a2 = (torch.randn(M, K, device="cuda") * 0.5).bfloat16()
    w = (torch.randn(N, K, device="cuda") / math.sqrt(K)).bfloat16()
    s = (a1.float() + a2.float()).bfloat16()
    o = capi.gemm(s, w).float()
    o12 = capi.gemm(a1, w).float() + capi.gemm(a2, w).float()
    assert rel(o, o12) < 8e-3
    idx = torch.randint(0, M, (32,), device="cuda")
    assert rel(o[idx], s[idx].float() @ w.float().t()) < BF16_ROUND_TOL


@pytest.mark.parametrize("d", [5120, 1536, 256])
def test_layernorm_and_rmsnorm_rope(capi, d):
    torch.manual_seed(3)
    rows = 333
    x = (torch.randn(rows, d, device="cuda") * 2 + 0.3).bfloat16()
    sh = (torch.randn(d, device="cuda") * 0.5).bfloat16()
    sc = (torch.randn(d, device="cuda") * 0.5).bfloat16()
    y = capi.layernorm(x, eps=1e-6, shift=sh, scale=sc)
    ref = F.layer_norm(x.float(), (d,), eps=1e-6).bfloat16() * (1 + sc) + sh      # eager bf16 chain of the reference
    assert rel(y, ref) < 1e-3 and (y != ref).float().mean() < 1e-2
    wt, bs = torch.randn(d, device="cuda").bfloat16(), torch.randn(d, device="cuda").bfloat16()
    y = capi.layernorm(x, eps=1e-6, weight=wt, bias=bs)
    assert rel(y, F.layer_norm(x.float(), (d,), wt.float(), bs.float(), eps=1e-6)) < BF16_ROUND_TOL
    heads = d // 128
    ang = torch.rand(rows, 64, device="cuda", dtype=torch.float64) * 6.28
    cs = torch.stack([ang.cos(), ang.sin()], -1).float().contiguous()
    xx = x.clone()
    capi.rmsnorm_rope_(xx, wt, eps=1e-6, cos_sin=cs, head_dim=128)
    xf = x.float()
    n = (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-6)).bfloat16() * wt
    c = torch.view_as_complex(n.double().reshape(rows, heads, 64, 2)) * torch.polar(torch.ones_like(ang), ang)[:, None]
    ref = torch.view_as_real(c).flatten(1).bfloat16()
    assert rel(xx, ref) < 1e-3 and (xx != ref).float().mean() < 1e-3     # fp32 table vs complex128: rare 1-ulp flips
    xx = x.clone()
    capi.rmsnorm_rope_(xx, wt, eps=1e-6, cos_sin=None, head_dim=128)
    assert (xx != n).float().mean() < 1e-3


@pytest.mark.parametrize("d", [5120, 1536])
def test_fused_qk_rmsnorm_rope_equals_two_calls(capi, d):
    torch.manual_seed(8)
    rows = 1030                                   # ragged against the 2 / 4 rows per CTA
    qkv = (torch.randn(rows, 3 * d, device="cuda") * 1.7).bfloat16()
    wq, wk = torch.randn(d, device="cuda").bfloat16(), torch.randn(d, device="cuda").bfloat16()
    ang = torch.rand(rows, 64, device="cuda", dtype=torch.float64) * 6.28
    cs = torch.stack([ang.cos(), ang.sin()], -1).float().contiguous()
    a = qkv.clone()
    capi.rmsnorm_rope_(a[:, :d], wq, eps=1e-6, cos_sin=cs, head_dim=128)
    capi.rmsnorm_rope_(a[:, d:2 * d], wk, eps=1e-6, cos_sin=cs, head_dim=128)
    b = qkv.clone()
    capi.qk_rmsnorm_rope_(b, wq, wk, eps=1e-6, cos_sin=cs, head_dim=128)
    assert torch.equal(a, b)
    assert torch.equal(b[:, 2 * d:], qkv[:, 2 * d:])          # v untouched


def _attn_ref(q, k, v, heads):
    Lq, Lk = q.shape[0], k.shape[0]
    qh, kh, vh = (t.float().reshape(t.shape[0], heads, 128).transpose(0, 1)[None] for t in (q, k, v))
    return F.scaled_dot_product_attention(qh, kh, vh)[0].transpose(0, 1).reshape(Lq, heads * 128)


@pytest.fixture(params=[(80, 0), (80, 4), (128, 4), (128, 0), (160, 0), (160, 4)], ids=lambda p: f"impl{p[0]}-emu{p[1]}")
def attn_variant(capi, request):
    """every attention kernel variant selectable through gf_ctx_set_attention; the default is restored afterwards"""
    capi.attention_tuning(*request.param)
    yield request.param
    capi.attention_tuning(0, -1)        # back to the per-shape default


@pytest.mark.parametrize("Lq,Lk,heads,amp", [(256, 128, 1, 1.0), (1, 7, 1, 1.0), (300, 200, 2, 1.0), (512, 1024, 3, 1.0),
                                              (256, 512, 1, 4.0), (1000, 1333, 2, 2.0), (130, 512, 12, 1.0),
                                              (257, 81, 1, 1.0), (640, 41, 2, 1.0), (700, 80, 1, 1.0),
                                              # 150 / 151 two-tile work items on 148 SMs: the tail items run as
                                              # single-tile CTAs (gf_attn80.cu tail splitting)
                                              (38400, 200, 1, 1.0), (38500, 333, 1, 1.0)])
def test_attention(capi, attn_variant, Lq, Lk, heads, amp):
    torch.manual_seed(4)
    q = (torch.randn(Lq, heads * 128, device="cuda") * amp).bfloat16()
    k = (torch.randn(Lk, heads * 128, device="cuda") * amp).bfloat16()
    v = torch.randn(Lk, heads * 128, device="cuda").bfloat16()
    assert rel(capi.attention(q, k, v, heads), _attn_ref(q, k, v, heads)) < 5e-3


def test_attention_rescale_path_and_strided_views(capi, attn_variant):
    # keys sorted by growing magnitude force the running max to jump by > 2^8 between kv blocks (lazy-rescale branch)
    torch.manual_seed(5)
    L, heads = 640, 2
    d = heads * 128
    qkv = torch.randn(L, 3 * d, device="cuda")
    qkv[:, d:2 * d] *= torch.linspace(0.2, 6.0, L, device="cuda")[:, None]
    qkv = qkv.bfloat16()
    q, k, v = qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:]
    assert rel(capi.attention(q, k, v, heads), _attn_ref(q, k, v, heads)) < 5e-3


def test_attention_softmax_properties_at_full_length(capi, attn_variant):
    # size-independent properties at L = 32760 (config 2), 2 heads:
    #   constant V  -> O == V exactly representable (softmax rows sum to 1);  permuting the keys leaves O unchanged
    torch.manual_seed(6)
    L, heads = 32760, 2
    d = heads * 128
    q = torch.randn(L, d, device="cuda").bfloat16()
    k = torch.randn(L, d, device="cuda").bfloat16()
    vconst = torch.full((L, d), 0.75, device="cuda").bfloat16()
    o = capi.attention(q, k, vconst, heads)
    assert float((o.float() - 0.75).abs().max()) < 8e-3
    v = torch.randn(L, d, device="cuda").bfloat16()
    perm = torch.randperm(L, device="cuda")
    o1 = capi.attention(q, k, v, heads)
    o2 = capi.attention(q, k[perm].contiguous(), v[perm].contiguous(), heads)
    assert rel(o1, o2) < 1e-2
    idx = torch.randint(0, L, (64,), device="cuda")
    ref = _attn_ref(q[idx], k, v, heads)
    assert rel(o1[idx], ref) < 5e-3


def test_index_kernels_bit_exact(capi):
    torch.manual_seed(7)
    Fr, H, W = 3, 8, 12
    a = torch.randn(16, Fr, H, W, device="cuda").bfloat16()
    b = torch.randn(20, Fr, H, W, device="cuda").bfloat16()
    x = torch.cat([a, b], 0)
    ref = x.reshape(36, Fr, H // 2, 2, W // 2, 2).permute(1, 2, 4, 0, 3, 5).reshape(-1, 144)
    assert torch.equal(capi.patch_gather(a, b), ref)
    assert torch.equal(capi.patch_gather(a, None), ref[:, :64])
    # against the conv itself
    wconv = torch.randn(256, 36, 1, 2, 2, device="cuda").bfloat16()
    conv = F.conv3d(x[None].float(), wconv.float(), stride=(1, 2, 2))[0].flatten(1).t()
    assert rel(ref.float() @ wconv.reshape(256, -1).float().t(), conv) < 1e-5
    L = Fr * (H // 2) * (W // 2)
    t = torch.randn(L, 64, device="cuda").bfloat16()
    ref = t.reshape(Fr, H // 2, W // 2, 2, 2, 16).permute(5, 0, 1, 3, 2, 4).reshape(16, Fr, H, W)
    assert torch.equal(capi.unpatchify(t, 16, Fr, H, W), ref)
    m = torch.randn(5, 6 * 256, device="cuda").bfloat16()
    tm = torch.randn(6 * 256, device="cuda").bfloat16()
    assert torch.equal(capi.add_rows(m, tm), m + tm)
    p, n, lat = (torch.randn(16, 5, 6, 8, device="cuda").bfloat16() for _ in range(3))
    ref = lat + (n + 5.0 * (p - n)) * torch.tensor(-0.0371)
    assert torch.equal(capi.cfg_euler(p, n, lat, 5.0, -0.0371), ref)
    assert torch.equal(capi.cfg_euler(p, None, lat, 1.0, -0.0371), lat + p * torch.tensor(-0.0371))
    xq = torch.randn(33, 8 * 128, device="cuda").bfloat16()
    pk = capi.ulysses_pack(xq, 8, 128, 4)
    assert torch.equal(pk, xq.view(33, 4, 2, 128).permute(1, 0, 2, 3).reshape(4, 33, 256))
    assert torch.equal(capi.ulysses_unpack(pk, 33, 8, 128, 4), xq)


def test_timestep_embedding(capi):
    from oracle import wan_dit_oracle as O
    for tval in (999.0, 937.3, 412.0, 3.5):
        t = torch.tensor([tval], dtype=torch.bfloat16, device="cuda")
        got = capi.timestep_embedding(t, 256)
        ref = O.sinusoidal_embedding_1d(256, t)
        assert (got != ref).float().mean() < 0.02 and float((got.float() - ref.float()).abs().max()) < 8e-3


def test_attention_kv_len_masks_padded_keys(capi, attn_variant):
    """kv_len < rows of K/V: the trailing (padding) keys are ignored -- same result as slicing them away."""
    torch.manual_seed(7)
    Lq, Lk, h = 333, 300, 2
    q = torch.randn(Lq, h * 128, device="cuda").bfloat16()
    k = torch.randn(Lk, h * 128, device="cuda").bfloat16()
    v = torch.randn(Lk, h * 128, device="cuda").bfloat16()
    for n in (299, 241, 80, 1):
        a = capi.attention(q, k, v, h, kv_len=n)
        b = capi.attention(q, k[:n], v[:n], h)
        assert torch.equal(a, b), n
    with pytest.raises(ValueError):
        capi.attention(q, k, v, h, kv_len=301)
