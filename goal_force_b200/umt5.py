"""umT5-XXL prompt encoder on the B200 kernels.

Mirrors `WanTextEncoder` (diffsynth/models/wan_video_text_encoder.py:178-245) and the post-processing of
`WanPrompter.encode_prompt` (diffsynth/prompters/wan_prompter.py:97-108): token embedding -> 24 x [T5LayerNorm,
self-attention with a per-layer relative-position bias (shared_pos=False), T5LayerNorm, gated-GELU FFN] -> final
T5LayerNorm, embeddings past the prompt length zeroed.  No biases anywhere, no 1/sqrt(d) in the attention.
State-dict keys are the reference's (`token_embedding.weight`, `blocks.N.{norm1,norm2}.weight`,
`blocks.N.attn.{q,k,v,o}.weight`, `blocks.N.ffn.{gate.0,fc1,fc2}.weight`, `blocks.N.pos_embedding.embedding.weight`,
`norm.weight`).  Tokenisation (HuggingFace tokenizer files) stays outside: callers pass `ids` and `mask` exactly as
`HuggingfaceTokenizer.__call__(..., return_mask=True)` returns them.

The encoder runs twice per video, ~5 TFLOP per prompt: the linears go through gf_gemm_bf16 (q|k|v fused, the GELU of
the gate fused into its GEMM, the residual adds fused into the o / fc2 GEMMs); both prompts of a CFG pair can be
encoded in one batch.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch

from . import capi


@dataclass(frozen=True)
class UMT5Config:
    """WanTextEncoder defaults (wan_video_text_encoder.py:180-189) = umT5-XXL encoder."""
    vocab: int = 256384
    dim: int = 4096
    dim_attn: int = 4096
    dim_ffn: int = 10240
    num_heads: int = 64
    num_layers: int = 24
    num_buckets: int = 32
    max_dist: int = 128
    eps: float = 1e-6


def relative_position_bucket(lq: int, lk: int, num_buckets: int, max_dist: int = 128) -> torch.Tensor:
    """bucket index for every relative position j - i in [-(lq-1), lk-1], int32 [lq + lk - 1]
    (T5RelativeEmbedding._relative_position_bucket, bidirectional, wan_video_text_encoder.py:154-175; same torch ops,
    so the float32 log and the truncation are the reference's)."""
    rel_pos = torch.arange(-(lq - 1), lk)
    nb = num_buckets // 2
    rel_buckets = (rel_pos > 0).long() * nb
    rel_pos = torch.abs(rel_pos)
    max_exact = nb // 2
    large = max_exact + (torch.log(rel_pos.float() / max_exact) / math.log(max_dist / max_exact)
                         * (nb - max_exact)).long()
    large = torch.min(large, torch.full_like(large, nb - 1))
    rel_buckets += torch.where(rel_pos < max_exact, rel_pos, large)
    return rel_buckets.to(torch.int32)


class _T5Block:
    __slots__ = ("norm1", "wqkv", "wo", "norm2", "wgate", "wfc1", "wfc2", "pos")


class UMT5EncoderB200:
    def __init__(self, cfg: UMT5Config, state_dict, device="cuda"):
        if cfg.dim_attn // cfg.num_heads != 64:
            raise ValueError("head_dim must be 64 (umT5)")
        if cfg.dim % 32 or cfg.dim_attn % 32 or cfg.dim_ffn % 32:
            raise ValueError("dim, dim_attn and dim_ffn must be multiples of 32")
        capi.load()
        self.cfg = cfg
        self.device = torch.device(device)
        sd = state_dict

        def g(name):
            return sd[name].detach().to(device=self.device, dtype=torch.bfloat16).contiguous()

        self.embedding = g("token_embedding.weight")
        self.blocks = []
        for i in range(cfg.num_layers):
            p = f"blocks.{i}."
            b = _T5Block()
            b.norm1, b.norm2 = g(p + "norm1.weight"), g(p + "norm2.weight")
            b.wqkv = torch.cat([g(p + "attn.q.weight"), g(p + "attn.k.weight"), g(p + "attn.v.weight")], 0).contiguous()
            b.wo = g(p + "attn.o.weight")
            b.wgate, b.wfc1, b.wfc2 = g(p + "ffn.gate.0.weight"), g(p + "ffn.fc1.weight"), g(p + "ffn.fc2.weight")
            b.pos = g(p + "pos_embedding.embedding.weight")                       # [num_buckets, heads]
            self.blocks.append(b)
        self.norm = g("norm.weight")
        self._buckets: dict = {}

    @classmethod
    def from_reference(cls, module, device="cuda") -> "UMT5EncoderB200":
        if getattr(module, "shared_pos", False):
            raise NotImplementedError("shared_pos=True is not the Wan text encoder configuration")
        blk = module.blocks[0]
        cfg = UMT5Config(vocab=module.token_embedding.num_embeddings, dim=module.dim, dim_attn=module.dim_attn,
                         dim_ffn=module.dim_ffn, num_heads=module.num_heads, num_layers=module.num_layers,
                         num_buckets=module.num_buckets, max_dist=blk.pos_embedding.max_dist, eps=module.norm.eps)
        return cls(cfg, module.state_dict(), device=device)

    def _bucket_table(self, L: int) -> torch.Tensor:
        t = self._buckets.get(L)
        if t is None:
            t = relative_position_bucket(L, L, self.cfg.num_buckets, self.cfg.max_dist).to(self.device)
            self._buckets[L] = t
        return t

    @torch.no_grad()
    def forward(self, ids: torch.Tensor, mask: torch.Tensor | None = None) -> torch.Tensor:
        """WanTextEncoder.forward (eval mode: dropout is the identity): ids [B, L] int64, mask [B, L] -> [B, L, dim]."""
        cfg = self.cfg
        B, L = ids.shape
        if L > 512:
            raise NotImplementedError("prompts longer than 512 tokens (the Wan prompter's text_len)")
        ids = ids.to(self.device, torch.int64).reshape(-1).contiguous()
        km = None if mask is None else mask.to(self.device).ne(0).to(torch.int32).contiguous()
        buckets = self._bucket_table(L)
        da = cfg.dim_attn
        x = capi.embedding(ids, self.embedding)                                  # [B*L, dim]
        h = torch.empty_like(x)
        qkv = torch.empty((B * L, 3 * da), dtype=torch.bfloat16, device=self.device)
        ao = torch.empty((B * L, da), dtype=torch.bfloat16, device=self.device)
        u = torch.empty((B * L, cfg.dim_ffn), dtype=torch.bfloat16, device=self.device)
        gt = torch.empty_like(u)
        for b in self.blocks:
            capi.t5_rmsnorm(x, b.norm1, eps=cfg.eps, out=h)
            capi.gemm(h, b.wqkv, None, out=qkv)
            capi.t5_attention(qkv[:, :da], qkv[:, da:2 * da], qkv[:, 2 * da:], batch=B, heads=cfg.num_heads,
                              bias_table=b.pos, bucket_of=buckets, key_mask=km, out=ao)
            capi.gemm(ao, b.wo, None, epi=capi.GF_EPI_GATE_RES, gate=None, residual=x, out=x)      # x + attn(...)
            capi.t5_rmsnorm(x, b.norm2, eps=cfg.eps, out=h)
            capi.gemm(h, b.wgate, None, epi=capi.GF_EPI_BIAS_GELU, out=gt)                         # gelu(gate(x))
            capi.gemm(h, b.wfc1, None, out=u)
            capi.mul_(u, gt)                                                                        # fc1(x) * gate(x)
            capi.gemm(u, b.wfc2, None, epi=capi.GF_EPI_GATE_RES, gate=None, residual=x, out=x)     # x + ffn(...)
        out = capi.t5_rmsnorm(x, self.norm, eps=cfg.eps)
        return out.view(B, L, cfg.dim)

    __call__ = forward

    @torch.no_grad()
    def encode_prompt(self, ids: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
        """WanPrompter.encode_prompt after tokenisation (wan_prompter.py:100-108): encoder output with every position
        from the prompt's own length on zeroed (the reference zeroes `prompt_emb[:, v:]` for each row's length v in
        turn, i.e. effectively from the SHORTEST length of the batch on; it is always called with one prompt)."""
        emb = self.forward(ids, mask)
        seq_lens = mask.to(self.device).gt(0).sum(dim=1).long()
        for v in seq_lens.tolist():
            emb[:, v:] = 0
        return emb
