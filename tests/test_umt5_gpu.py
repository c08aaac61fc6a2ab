"""umT5 prompt encoder on the GPU kernels against the oracle (reference restatement) on the same device."""
import pytest
import torch

from oracle import umt5_oracle as U
from oracle import wan_dit_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _no_tf32(lib):
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    yield
    torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old


def _check(out, ref32, refbf):
    e_ours, e_ref = O.rel_l2(out, ref32), O.rel_l2(refbf, ref32)
    print(f"umT5 relL2 ours-vs-fp32 {e_ours:.3e}  ref_bf16-vs-fp32 {e_ref:.3e}  ours-vs-ref_bf16 {O.rel_l2(out, refbf):.3e}")
    assert not torch.isnan(out).any()
    assert e_ours <= max(1e-2, 1.0 * e_ref), (e_ours, e_ref)


def test_encoder_matches_reference_golden(golden_dir):
    from goal_force_b200.umt5 import UMT5Config, UMT5EncoderB200
    g = torch.load(golden_dir / "umt5.pt", weights_only=False)["tiny"]
    c = g["cfg"]
    sd = U.random_state_dict(seed=g["weight_seed"], **c)
    ids, mask = U.synthetic_prompt(c["vocab"], g["batch"], g["L"], g["valid"], seed=g["prompt_seed"])
    enc = UMT5EncoderB200(UMT5Config(**c), sd)
    out = enc(ids.cuda(), mask.cuda())
    assert out.shape == g["out_fp32"].shape and out.dtype == torch.bfloat16
    _check(out.cpu(), g["out_fp32"], g["out_bf16"])
    # prompter post-processing: everything from the (shortest) prompt length on is zero
    emb = enc.encode_prompt(ids[1:].cuda(), mask[1:].cuda())
    v = int(mask[1].sum())
    assert float(emb[:, v:].abs().max()) == 0.0 and torch.equal(emb[0, :v], out[1, :v])


def test_encoder_umt5_xxl_width_two_layers_vs_oracle():
    """umT5-XXL widths (dim 4096, 64 heads x 64, ffn 10240, 32 buckets), 2 layers, two 512-token prompts (one full,
    one 77 tokens + padding), against the oracle in fp32 and bf16 on this device."""
    from goal_force_b200.umt5 import UMT5Config, UMT5EncoderB200
    c = dict(vocab=2048, dim=4096, dim_attn=4096, dim_ffn=10240, num_heads=64, num_layers=2, num_buckets=32)
    sd = U.random_state_dict(seed=3, **c)
    ids, mask = U.synthetic_prompt(c["vocab"], 2, 512, (512, 77), seed=4)
    kw = dict(num_heads=64, num_layers=2, num_buckets=32)
    refs = []
    for dt in (torch.float32, torch.bfloat16):
        s = {k: v.to("cuda", dt) for k, v in sd.items()}
        with torch.no_grad():
            refs.append(U.encoder(s, ids.cuda(), mask.cuda(), **kw))
        del s
    out = UMT5EncoderB200(UMT5Config(**c), sd)(ids.cuda(), mask.cuda())
    _check(out, refs[0], refs[1])


def test_t5_attention_kernel_vs_torch(capi):
    """gf_t5_attention_bf16 alone: bias table + padding mask + ragged lengths against the reference formula in fp32."""
    torch.manual_seed(0)
    from goal_force_b200.umt5 import relative_position_bucket
    for (B, L, H, valid) in ((1, 512, 3, (512,)), (2, 77, 2, (77, 30)), (1, 33, 1, (20,))):
        q = (torch.randn(B * L, H * 64, device="cuda") * 0.5).bfloat16()
        k = torch.randn(B * L, H * 64, device="cuda").bfloat16()
        v = torch.randn(B * L, H * 64, device="cuda").bfloat16()
        table = (torch.randn(32, H, device="cuda") * 0.7).bfloat16()
        buckets = relative_position_bucket(L, L, 32).cuda()
        mask = torch.zeros(B, L, dtype=torch.int32, device="cuda")
        for b, n in enumerate(valid):
            mask[b, :n] = 1
        out = capi.t5_attention(q, k, v, batch=B, heads=H, bias_table=table, bucket_of=buckets, key_mask=mask)
        i = torch.arange(L, device="cuda").unsqueeze(1)
        j = torch.arange(L, device="cuda").unsqueeze(0)
        bias = table.float()[buckets.long()[j - i + L - 1]].permute(2, 0, 1)                  # [H, L, L]
        qh, kh, vh = (t.float().view(B, L, H, 64).permute(0, 2, 1, 3) for t in (q, k, v))
        s = qh @ kh.transpose(-1, -2) + bias
        s = s.masked_fill(mask.view(B, 1, 1, L) == 0, torch.finfo(torch.bfloat16).min)
        ref = (torch.softmax(s, -1) @ vh).permute(0, 2, 1, 3).reshape(B * L, H * 64)
        err = O.rel_l2(out, ref)
        print(f"t5 attention B{B} L{L} H{H}: relL2 {err:.3e}")
        assert err < 6e-3
