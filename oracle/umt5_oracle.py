"""torch restatement of the umT5 prompt encoder -- TEST INFRASTRUCTURE ONLY (see oracle/wan_dit_oracle.py).

Follows diffsynth/models/wan_video_text_encoder.py (T5LayerNorm :18-30, T5Attention :47-78, T5FeedForward :95-100,
T5SelfAttention :128-132, T5RelativeEmbedding :136-175, WanTextEncoder.forward :235-245) op for op on whatever
device / dtype the weights have, so it is bit-identical to the reference module in eval mode; pinned by
tests/golden/umt5.pt, which oracle/gen_golden.py produces by running the reference class itself.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def t5_layer_norm(x, weight, eps=1e-6):
    x = x * torch.rsqrt(x.float().pow(2).mean(dim=-1, keepdim=True) + eps)
    if weight.dtype in (torch.float16, torch.bfloat16):
        x = x.type_as(weight)
    return weight * x


def gelu_tanh(x):
    return 0.5 * x * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * torch.pow(x, 3.0))))


def relative_bucket(rel_pos, num_buckets, max_dist=128):
    nb = num_buckets // 2
    rel_buckets = (rel_pos > 0).long() * nb
    rel_pos = torch.abs(rel_pos)
    max_exact = nb // 2
    large = max_exact + (torch.log(rel_pos.float() / max_exact) / math.log(max_dist / max_exact) * (nb - max_exact)).long()
    large = torch.min(large, torch.full_like(large, nb - 1))
    rel_buckets += torch.where(rel_pos < max_exact, rel_pos, large)
    return rel_buckets


def position_bias(emb_weight, lq, lk, num_buckets, max_dist=128):
    dev = emb_weight.device
    rel = torch.arange(lk, device=dev).unsqueeze(0) - torch.arange(lq, device=dev).unsqueeze(1)
    e = F.embedding(relative_bucket(rel, num_buckets, max_dist), emb_weight)
    return e.permute(2, 0, 1).unsqueeze(0).contiguous()          # [1, N, Lq, Lk]


def attention(sd, pre, x, num_heads, mask, pos_bias):
    b, n = x.size(0), num_heads
    q = F.linear(x, sd[pre + "q.weight"])
    c = q.shape[-1] // n
    q = q.view(b, -1, n, c)
    k = F.linear(x, sd[pre + "k.weight"]).view(b, -1, n, c)
    v = F.linear(x, sd[pre + "v.weight"]).view(b, -1, n, c)
    attn_bias = x.new_zeros(b, n, q.size(1), k.size(1))
    attn_bias += pos_bias
    if mask is not None:
        m = mask.view(b, 1, 1, -1) if mask.ndim == 2 else mask.unsqueeze(1)
        attn_bias.masked_fill_(m == 0, torch.finfo(x.dtype).min)
    attn = torch.einsum("binc,bjnc->bnij", q, k) + attn_bias
    attn = F.softmax(attn.float(), dim=-1).type_as(attn)
    x = torch.einsum("bnij,bjnc->binc", attn, v).reshape(b, -1, n * c)
    return F.linear(x, sd[pre + "o.weight"])


def encoder(sd, ids, mask, *, num_heads, num_layers, num_buckets, max_dist=128, eps=1e-6):
    x = F.embedding(ids, sd["token_embedding.weight"])
    L = x.size(1)
    for i in range(num_layers):
        p = f"blocks.{i}."
        e = position_bias(sd[p + "pos_embedding.embedding.weight"], L, L, num_buckets, max_dist)
        x = x + attention(sd, p + "attn.", t5_layer_norm(x, sd[p + "norm1.weight"], eps), num_heads, mask, e)
        h = t5_layer_norm(x, sd[p + "norm2.weight"], eps)
        ff = F.linear(h, sd[p + "ffn.fc1.weight"]) * gelu_tanh(F.linear(h, sd[p + "ffn.gate.0.weight"]))
        x = x + F.linear(ff, sd[p + "ffn.fc2.weight"])
    return t5_layer_norm(x, sd["norm.weight"], eps)


def random_state_dict(vocab, dim, dim_attn, dim_ffn, num_heads, num_layers, num_buckets, seed=0, dtype=torch.float32,
                      device="cpu"):
    """Seeded weights with the reference's key names; std as in init_weights (:177-193) except the q projection and
    the position table, which get larger values so that the attention is not numerically flat."""
    g = torch.Generator("cpu").manual_seed(seed)
    rn = lambda *s, std=1.0: torch.randn(*s, generator=g) * std  # noqa: E731
    sd = {"token_embedding.weight": rn(vocab, dim), "norm.weight": 1.0 + rn(dim, std=0.1)}
    for i in range(num_layers):
        p = f"blocks.{i}."
        sd[p + "norm1.weight"] = 1.0 + rn(dim, std=0.1)
        sd[p + "norm2.weight"] = 1.0 + rn(dim, std=0.1)
        sd[p + "attn.q.weight"] = rn(dim_attn, dim, std=dim ** -0.5 * 0.35)
        sd[p + "attn.k.weight"] = rn(dim_attn, dim, std=dim ** -0.5)
        sd[p + "attn.v.weight"] = rn(dim_attn, dim, std=dim ** -0.5)
        sd[p + "attn.o.weight"] = rn(dim, dim_attn, std=dim_attn ** -0.5)
        sd[p + "ffn.gate.0.weight"] = rn(dim_ffn, dim, std=dim ** -0.5)
        sd[p + "ffn.fc1.weight"] = rn(dim_ffn, dim, std=dim ** -0.5)
        sd[p + "ffn.fc2.weight"] = rn(dim, dim_ffn, std=dim_ffn ** -0.5)
        sd[p + "pos_embedding.embedding.weight"] = rn(num_buckets, num_heads, std=0.5)
    return {k: v.to(dtype=dtype, device=device) for k, v in sd.items()}


def synthetic_prompt(vocab, batch, L, valid, seed=1):
    g = torch.Generator("cpu").manual_seed(seed)
    ids = torch.randint(1, vocab, (batch, L), generator=g)
    mask = torch.zeros(batch, L, dtype=torch.long)
    for b in range(batch):
        mask[b, :valid[b]] = 1
        ids[b, valid[b]:] = 0
    return ids, mask
