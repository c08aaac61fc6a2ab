"""B200-native Wan DiT expert and goal-force ControlNet behind the reference's call signatures.

Mirrors (names, argument meaning, error behaviour):
  * WanModel.forward(x, timestep, context, clip_feature=None, y=None, **kw)   diffsynth/models/wan_video_dit.py:358-411
  * model_fn_wan_video(dit=, latents=, timestep=, context=, y=, controlnet=, control_signal_video_latents=, **kw)
                                                                           src/goal_force/wan_video_new.py:1349-1591
  * ControlNet(num_layers, stride)                                         src/goal_force/wan_video_new.py:97-117
Weights come in as a reference state_dict (same key names); linear weights are re-laid-out once (q|k|v fused, the
cross-attention k|v fused, conv weights flattened to GEMM operands). All compute runs in libgoalforce_b200.so through
goal_force_b200.capi; torch only owns memory and streams. There is no CPU / eager fallback.
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass

import torch

from . import capi


@dataclass(frozen=True)
class DiTConfig:
    """kwargs tables of the reference: wan_video_dit.py:502-514 (Wan2.1 T2V 1.3B), :703-718 (Wan2.2 I2V A14B)."""
    dim: int
    in_dim: int
    ffn_dim: int
    out_dim: int
    text_dim: int
    freq_dim: int
    eps: float
    num_heads: int
    num_layers: int
    patch_size: tuple = (1, 2, 2)

    @property
    def head_dim(self) -> int:
        return self.dim // self.num_heads


WAN21_T2V_1_3B = DiTConfig(dim=1536, in_dim=16, ffn_dim=8960, out_dim=16, text_dim=4096, freq_dim=256, eps=1e-6,
                           num_heads=12, num_layers=30)
WAN22_I2V_A14B = DiTConfig(dim=5120, in_dim=36, ffn_dim=13824, out_dim=16, text_dim=4096, freq_dim=256, eps=1e-6,
                           num_heads=40, num_layers=40)


def _check_cfg(cfg: DiTConfig) -> None:
    if cfg.head_dim != 128:
        raise ValueError(f"head_dim must be 128 (got {cfg.head_dim}): the attention kernel is built for Wan's 128")
    if tuple(cfg.patch_size) != (1, 2, 2):
        raise ValueError("patch_size must be (1, 2, 2)")
    if cfg.dim % 256 or cfg.ffn_dim % 32:
        raise ValueError("dim must be a multiple of 256 and ffn_dim of 32")


def rope_cos_sin(head_dim: int, f: int, h: int, w: int, device, token_slice: slice | None = None,
                 pad_rows_to: int | None = None) -> torch.Tensor:
    """(cos, sin) table [f*h*w, head_dim/2, 2] fp32 for the kernels, from the reference's float64 recipe
    (precompute_freqs_cis_3d, wan_video_dit.py:75-89, assembled as in :380-384): per-axis angles pos * theta^(-2j/dim)
    with dims head_dim-2*(head_dim//3), head_dim//3, head_dim//3; cos/sin taken in float64, rounded once to fp32."""
    third = head_dim // 3
    dims = (head_dim - 2 * third, third, third)

    def axis(dim: int, n: int) -> torch.Tensor:
        inv = 1.0 / (10000.0 ** (torch.arange(0, dim, 2)[: dim // 2].double() / dim))
        return torch.outer(torch.arange(n).double(), inv)                       # (n, dim/2) float64 angles

    af, ah, aw = axis(dims[0], f), axis(dims[1], h), axis(dims[2], w)
    ang = torch.cat([
        af.view(f, 1, 1, -1).expand(f, h, w, -1),
        ah.view(1, h, 1, -1).expand(f, h, w, -1),
        aw.view(1, 1, w, -1).expand(f, h, w, -1),
    ], dim=-1).reshape(f * h * w, head_dim // 2)
    if token_slice is not None:
        ang = ang[token_slice]
    if pad_rows_to is not None and ang.shape[0] < pad_rows_to:          # identity rotation for padding rows
        ang = torch.cat([ang, ang.new_zeros(pad_rows_to - ang.shape[0], ang.shape[1])], 0)
    return torch.stack([torch.cos(ang), torch.sin(ang)], dim=-1).float().contiguous().to(device)


class _Block:
    """Weights of one DiTBlock (wan_video_dit.py:196-212) in kernel layout."""
    __slots__ = ("wqkv", "bqkv", "norm_q", "norm_k", "wo", "bo", "cq_w", "cq_b", "cnorm_q", "cnorm_k", "ckv_w",
                 "ckv_b", "co_w", "co_b", "n3w", "n3b", "w1", "b1", "w2", "b2", "modulation")

    @classmethod
    def from_state_dict(cls, sd: dict, pre: str, device) -> "_Block":
        def g(name):
            return sd[pre + name].detach().to(device=device, dtype=torch.bfloat16).contiguous()

        b = cls()
        sa, ca = "self_attn.", "cross_attn."
        b.wqkv = torch.cat([g(sa + "q.weight"), g(sa + "k.weight"), g(sa + "v.weight")], 0).contiguous()
        b.bqkv = torch.cat([g(sa + "q.bias"), g(sa + "k.bias"), g(sa + "v.bias")], 0).contiguous()
        b.norm_q, b.norm_k = g(sa + "norm_q.weight"), g(sa + "norm_k.weight")
        b.wo, b.bo = g(sa + "o.weight"), g(sa + "o.bias")
        b.cq_w, b.cq_b = g(ca + "q.weight"), g(ca + "q.bias")
        b.cnorm_q, b.cnorm_k = g(ca + "norm_q.weight"), g(ca + "norm_k.weight")
        b.ckv_w = torch.cat([g(ca + "k.weight"), g(ca + "v.weight")], 0).contiguous()
        b.ckv_b = torch.cat([g(ca + "k.bias"), g(ca + "v.bias")], 0).contiguous()
        b.co_w, b.co_b = g(ca + "o.weight"), g(ca + "o.bias")
        b.n3w, b.n3b = g("norm3.weight"), g("norm3.bias")
        b.w1, b.b1 = g("ffn.0.weight"), g("ffn.0.bias")
        b.w2, b.b2 = g("ffn.2.weight"), g("ffn.2.bias")
        b.modulation = g("modulation").reshape(-1)          # (6*dim,)
        return b


class _Workspace:
    """Activation buffers for one token count; reused across blocks, steps and experts."""

    def __init__(self, L: int, cfg: DiTConfig, device):
        d = cfg.dim
        bf = dict(dtype=torch.bfloat16, device=device)
        self.L = L
        self.h = torch.empty((L, d), **bf)
        self.qkv = torch.empty((L, 3 * d), **bf)
        self.ao = torch.empty((L, d), **bf)
        self.qc = torch.empty((L, d), **bf)
        self.u = torch.empty((L, cfg.ffn_dim), **bf)
        self._cn_states = None

    def cn_states(self, n: int) -> torch.Tensor:
        """[n, L, dim] outputs of the ControlNet blocks (a17); block i reads state i-1 and writes state i, so the
        branch needs no copies.  Allocated on first use (3.4 GB for 10 layers at 32,760 tokens)."""
        if self._cn_states is None or self._cn_states.shape[0] < n:
            self._cn_states = torch.empty((n, self.L, self.h.shape[1]), dtype=torch.bfloat16, device=self.h.device)
        return self._cn_states[:n]


_WORKSPACES: dict = {}


def _workspace(L: int, cfg: DiTConfig, device, slot: int = 0) -> _Workspace:
    """slot 0: trunk (and everything on the caller's stream); slot 1: the ControlNet branch when it runs on its own
    stream next to the trunk."""
    key = (L, cfg.dim, cfg.ffn_dim, str(device), slot)
    ws = _WORKSPACES.get(key)
    if ws is None:
        if len(_WORKSPACES) > 6:
            _WORKSPACES.clear()
        ws = _Workspace(L, cfg, device)
        _WORKSPACES[key] = ws
    return ws


class PeerExchange:
    """Peer-memory buffers of one rank for the fused Ulysses exchange (include/goalforce_b200.h, gf_peer_*):
      recv [P*Ll, 3w]  q|k|v of this rank's heads for ALL tokens, written by every rank's RMSNorm+RoPE kernel
      ao   [Ll, d]     attention output of ALL heads for this rank's tokens, written by every rank's attention kernel
      flags            barrier slots
    The buffers are allocated by the library (cudaMalloc) and mapped into the peers with CUDA IPC; the handles travel
    through torch.distributed once per (token count, width)."""

    def __init__(self, dist, group, rank: int, size: int, Ll: int, d: int, device):
        if size > 8:
            raise ValueError("peer exchange is limited to the 8 GPUs of one NVSwitch domain")
        self.dist, self.group = dist, group
        self.rank, self.size, self.Ll, self.d = rank, size, Ll, d
        self.w = d // size
        self.timeout_ms = int(os.environ.get("GF_PEER_TIMEOUT_MS", "60000"))   # read once, not on the launch path
        sizes = {"recv": size * Ll * 3 * self.w * 2, "ao": Ll * d * 2, "flags": 256}
        self.local = {k: capi.peer_alloc(n) for k, n in sizes.items()}
        self.status = capi.PeerStatus()
        mine = {k: capi.peer_export(p) for k, p in self.local.items()}
        everyone = [None] * size
        dist.all_gather_object(everyone, mine, group=group)
        self._imported = []
        ptrs = {k: [] for k in sizes}
        for r in range(size):
            for k in sizes:
                if r == rank:
                    ptrs[k].append(self.local[k])
                else:
                    p = capi.peer_import(everyone[r][k])
                    self._imported.append(p)
                    ptrs[k].append(p)
        self.recv_ptrs, self.ao_ptrs, self.flag_ptrs = (capi.ptr_array(ptrs[k]) for k in ("recv", "ao", "flags"))
        self.recv = capi.as_bf16_tensor(self.local["recv"], (size * Ll, 3 * self.w), device)
        self.ao = capi.as_bf16_tensor(self.local["ao"], (Ll, d), device)
        dist.barrier(group=group)          # every rank has mapped every buffer before anybody writes
        _LIVE_EXCHANGES.append(self)

    def barrier(self) -> None:
        """All ranks' earlier kernels (on their compute streams) are complete and visible when this returns on the
        stream; the only collective left in the exchange.  The epoch counter lives in the flag buffer."""
        capi.peer_barrier(self.flag_ptrs, self.size, self.rank, self.timeout_ms, self.status)

    def check(self) -> None:
        """Raise if a barrier of an earlier forward timed out (reads a host-mapped word; no stream sync)."""
        self.status.check()

    def close(self) -> None:
        """Collective teardown: drain my stream, unmap the peers' buffers, wait until every rank has unmapped mine
        (cudaFree of memory a peer still has IPC-mapped is undefined), then free."""
        if not self.local:
            return
        torch.cuda.synchronize()
        for p in self._imported:
            capi.peer_unimport(p)
        self._imported = []
        try:
            self.dist.barrier(group=self.group)
        except Exception:  # noqa: BLE001 - process group already destroyed at interpreter exit
            pass
        for p in self.local.values():
            capi.peer_free(p)
        self.local = {}
        self.status.free()
        if self in _LIVE_EXCHANGES:
            _LIVE_EXCHANGES.remove(self)


_LIVE_EXCHANGES: list = []      # creation order (identical on every rank, so teardown barriers pair up)


def close_peer_exchanges() -> None:
    """Tear down every live PeerExchange (collective: call on all ranks, before destroy_process_group)."""
    for ex in list(_LIVE_EXCHANGES):
        ex.close()


class SequenceParallel:
    """Ulysses sequence parallelism over one torch.distributed group (replaces
    diffsynth/distributed/xdit_context_parallel.py:42-131 and the xfuser dependency).
    Tokens are split contiguously over ranks; per self-attention q,k,v move from [L/P, heads] to [L, heads/P] and the
    output moves back.  Two transports:
      "peer": fused -- the RMSNorm+RoPE kernel stores q|k|v directly into the head owners' buffers over NVLink and the
              attention kernel stores its output rows directly into the token owners' buffers; two flag barriers per
              attention replace the two all-to-alls (no pack / unpack kernels, no NCCL on the path);
      "nccl": pack kernel -> all_to_all_single -> attention -> all_to_all_single -> unpack kernel (baseline)."""

    def __init__(self, group=None, transport: str = "nccl"):
        import torch.distributed as dist
        if transport not in ("peer", "nccl"):
            raise ValueError("transport must be 'peer' or 'nccl'")
        self.dist = dist
        self.group = group
        self.size = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.transport = transport
        self._bufs: dict = {}
        self._peer: dict = {}

    def rows_per_rank(self, L: int) -> int:
        return -(-L // self.size)

    def token_slice(self, L: int) -> slice:
        """This rank's contiguous token range.  When L does not divide, the ranges are ceil(L/P) long and the tail of
        the last ranks is padding (the reference pads the chunked sequence with zeros as well,
        diffsynth/distributed/xdit_context_parallel.py:15-40,60-66; there the pad rows are attended as keys, here
        they are masked out, so SP(P) stays identical to the unsharded forward)."""
        n = self.rows_per_rank(L)
        return slice(min(self.rank * n, L), min((self.rank + 1) * n, L))

    def shard_rows(self, t: torch.Tensor, L: int) -> torch.Tensor:
        """rows of this rank out of a [L, ...] tensor, zero-padded to rows_per_rank(L)."""
        n = self.rows_per_rank(L)
        part = t[self.token_slice(L)]
        if part.shape[0] == n:
            return part.contiguous()
        out = t.new_zeros((n,) + tuple(t.shape[1:]))
        out[:part.shape[0]] = part
        return out

    def buffers(self, Ll: int, d: int, device, slot: int = 0):
        key = (Ll, d, str(device), slot)
        b = self._bufs.get(key)
        if b is None:
            P = self.size
            bf = dict(dtype=torch.bfloat16, device=device)
            b = (torch.empty((P, Ll, 3 * d // P), **bf), torch.empty((P * Ll, 3 * d // P), **bf),
                 torch.empty((P * Ll, d // P), **bf), torch.empty((P, Ll, d // P), **bf))
            self._bufs = {k: v for k, v in self._bufs.items() if k[:3] == key[:3]}
            self._bufs[key] = b
        return b

    def peer_exchange(self, Ll: int, d: int, device, slot: int = 0) -> PeerExchange:
        """One set of peer buffers per (token count, stream slot): the trunk and a ControlNet branch running next to it
        on a second stream exchange through separate buffers and flags."""
        key = (Ll, d, str(device), slot)
        ex = self._peer.get(key)
        if ex is None:
            for k in [k for k in self._peer if k[:3] != key[:3]]:      # token count changed: tear the old ones down
                self._peer.pop(k).close()
            ex = PeerExchange(self.dist, self.group, self.rank, self.size, Ll, d, device)
            self._peer[key] = ex
        return ex

    def check(self) -> None:
        for ex in self._peer.values():
            ex.check()

    def close(self) -> None:
        for ex in self._peer.values():
            ex.close()
        self._peer = {}

    def _check_heads(self, heads: int) -> int:
        if heads % self.size:
            raise ValueError(f"{heads} heads do not divide over {self.size} ranks")
        return heads // self.size

    def self_attention(self, qkv: torch.Tensor, heads: int, out: torch.Tensor, slot: int = 0,
                       kv_len: int | None = None) -> None:
        """NCCL transport. qkv: [L/P, 3*d] (q|k|v, already normed + roped) -> out [L/P, d]."""
        P, Ll = self.size, qkv.shape[0]
        d = qkv.shape[1] // 3
        hp = self._check_heads(heads)
        send, recv, o_full, o_recv = self.buffers(Ll, d, qkv.device, slot)
        w = hp * 128
        for s in range(3):
            capi.ulysses_pack(qkv[:, s * d:(s + 1) * d], heads, 128, P, out=send[:, :, s * w:], out_pitch=3 * w)
        self.dist.all_to_all_single(recv.view(P, Ll, 3 * w), send, group=self.group)
        capi.attention(recv[:, :w], recv[:, w:2 * w], recv[:, 2 * w:], hp, out=o_full, kv_len=kv_len)
        self.dist.all_to_all_single(o_recv, o_full.view(P, Ll, w), group=self.group)
        capi.ulysses_unpack(o_recv, Ll, heads, 128, P, out=out)

    def self_attention_fused(self, qkv: torch.Tensor, norm_q: torch.Tensor, norm_k: torch.Tensor, eps: float,
                             cos_sin: torch.Tensor, heads: int, slot: int = 0, kv_len: int | None = None) -> torch.Tensor:
        """Peer transport. qkv: [L/P, 3*d] straight out of the QKV GEMM (NOT yet normed). Returns the [L/P, d]
        attention output (a view of the peer-visible buffer, valid until the next call)."""
        P, Ll = self.size, qkv.shape[0]
        d = qkv.shape[1] // 3
        hp = self._check_heads(heads)
        ex = self.peer_exchange(Ll, d, qkv.device, slot)
        w = ex.w
        capi.qkv_rmsnorm_rope_scatter(qkv, norm_q, norm_k, eps=eps, cos_sin=cos_sin, head_dim=128,
                                      recv_ptrs=ex.recv_ptrs, n_peers=P, rank=self.rank, ld_recv=3 * w)
        ex.barrier()                                   # every rank's q|k|v has landed here
        r = ex.recv
        capi.attention_scatter(r[:, :w], r[:, w:2 * w], r[:, 2 * w:], hp, out_ptrs=ex.ao_ptrs, n_peers=P, ldo=d,
                               rows_per_peer=Ll, col_offset=self.rank * w, kv_len=kv_len)
        ex.barrier()                                   # every rank's heads have landed in my ao
        return ex.ao


def run_block(bw: _Block, cfg: DiTConfig, x: torch.Tensor, ctx_kv: torch.Tensor, mod: torch.Tensor,
              cos_sin: torch.Tensor, ws: _Workspace, sp: SequenceParallel | None = None,
              x_in: torch.Tensor | None = None, slot: int = 0, kv_len: int | None = None) -> None:
    """DiTBlock.forward (wan_video_dit.py:214-230) on [L, dim] tokens: in place on x, or from x_in into x (x_in is
    left untouched: the first residual GEMM reads it and writes x).
    mod: [6, dim] = modulation + t_mod (shift_msa, scale_msa, gate_msa, shift_mlp, scale_mlp, gate_mlp);
    ctx_kv: [ctx_len, 2*dim] cached cross-attention keys (RMS-normed) | values of this block."""
    d, H, eps = cfg.dim, cfg.num_heads, cfg.eps
    L = x.shape[0]
    h, qkv, ao, qc, u = ws.h[:L], ws.qkv[:L], ws.ao[:L], ws.qc[:L], ws.u[:L]
    src = x if x_in is None else x_in
    # --- self attention
    capi.layernorm(src, eps=eps, shift=mod[0], scale=mod[1], out=h)
    capi.gemm(h, bw.wqkv, bw.bqkv, out=qkv)
    if sp is not None and sp.size > 1 and sp.transport == "peer":
        attn = sp.self_attention_fused(qkv, bw.norm_q, bw.norm_k, eps, cos_sin, H, slot, kv_len)
    else:
        capi.qk_rmsnorm_rope_(qkv, bw.norm_q, bw.norm_k, eps=eps, cos_sin=cos_sin, head_dim=128)
        if sp is None or sp.size == 1:
            capi.attention(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], H, out=ao)
        else:
            sp.self_attention(qkv, H, ao, slot, kv_len)
        attn = ao
    capi.gemm(attn, bw.wo, bw.bo, epi=capi.GF_EPI_GATE_RES, gate=mod[2], residual=src, out=x)
    # --- cross attention
    capi.layernorm(x, eps=eps, weight=bw.n3w, bias=bw.n3b, out=h)
    capi.gemm(h, bw.cq_w, bw.cq_b, out=qc)
    capi.rmsnorm_rope_(qc, bw.cnorm_q, eps=eps, cos_sin=None, head_dim=128)
    capi.attention(qc, ctx_kv[:, :d], ctx_kv[:, d:], H, out=ao)
    capi.gemm(ao, bw.co_w, bw.co_b, epi=capi.GF_EPI_GATE_RES, gate=None, residual=x, out=x)
    # --- ffn
    capi.layernorm(x, eps=eps, shift=mod[3], scale=mod[4], out=h)
    capi.gemm(h, bw.w1, bw.b1, epi=capi.GF_EPI_BIAS_GELU, out=u)
    capi.gemm(u, bw.w2, bw.b2, epi=capi.GF_EPI_GATE_RES, gate=mod[5], residual=x, out=x)


def block_context_kv(bw: _Block, cfg: DiTConfig, ctx_emb: torch.Tensor) -> torch.Tensor:
    """Cross-attention k = RMSNorm(Wk ctx + b) * w and v = Wv ctx + b (wan_video_dit.py:178-179) for one block.
    Depends only on the prompt and the expert, not on the denoising step."""
    kv = capi.gemm(ctx_emb, bw.ckv_w, bw.ckv_b)
    capi.rmsnorm_rope_(kv[:, :cfg.dim], bw.cnorm_k, eps=cfg.eps, cos_sin=None, head_dim=128)
    return kv


class WanModelB200:
    """Drop-in for the reference WanModel on the denoising path (inference, bf16, has_image_input=False)."""

    def __init__(self, cfg: DiTConfig, state_dict: dict, device="cuda"):
        _check_cfg(cfg)
        capi.load()
        self.cfg = cfg
        self.device = torch.device(device)
        # attribute surface the reference pipeline reads
        self.dim, self.in_dim, self.freq_dim, self.patch_size = cfg.dim, cfg.in_dim, cfg.freq_dim, cfg.patch_size
        self.has_image_input = False
        self.require_vae_embedding = True
        self.require_clip_embedding = False
        self.seperated_timestep = False
        self.fuse_vae_embedding_in_latents = False
        self.has_image_pos_emb = False
        self.control_adapter = None
        sd = state_dict

        def g(name):
            return sd[name].detach().to(device=self.device, dtype=torch.bfloat16).contiguous()

        self.patch_w = g("patch_embedding.weight").reshape(cfg.dim, -1).contiguous()     # (dim, in_dim*4)
        if self.patch_w.shape[1] != cfg.in_dim * 4:
            raise ValueError("patch_embedding.weight does not match in_dim")
        self.patch_b = g("patch_embedding.bias")
        self.text0_w, self.text0_b = g("text_embedding.0.weight"), g("text_embedding.0.bias")
        self.text2_w, self.text2_b = g("text_embedding.2.weight"), g("text_embedding.2.bias")
        self.time0_w, self.time0_b = g("time_embedding.0.weight"), g("time_embedding.0.bias")
        self.time2_w, self.time2_b = g("time_embedding.2.weight"), g("time_embedding.2.bias")
        self.tproj_w, self.tproj_b = g("time_projection.1.weight"), g("time_projection.1.bias")
        self.blocks = [_Block.from_state_dict(sd, f"blocks.{i}.", self.device) for i in range(cfg.num_layers)]
        self.block_mod = torch.stack([b.modulation for b in self.blocks], 0).contiguous()   # (layers, 6*dim)
        self.head_w, self.head_b = g("head.head.weight"), g("head.head.bias")
        self.head_mod = g("head.modulation").reshape(2, cfg.dim).contiguous()
        self._ctx_cache: dict = {}
        self._rope_cache: dict = {}

    # -------------------------------------------------------------------------------------------- construction
    @classmethod
    def from_reference(cls, module, device="cuda") -> "WanModelB200":
        """Build from a reference diffsynth WanModel instance (weights are copied to bf16 kernel layout)."""
        if getattr(module, "has_image_input", False):
            raise NotImplementedError("has_image_input=True (CLIP image branch) is outside the goal-force hot path")
        for flag in ("seperated_timestep", "fuse_vae_embedding_in_latents", "has_ref_conv", "has_image_pos_emb"):
            if getattr(module, flag, False):
                raise NotImplementedError(f"WanModel variant with {flag}=True is outside the goal-force hot path")
        if getattr(module, "control_adapter", None) is not None:
            raise NotImplementedError("WanModel with a control adapter is outside the goal-force hot path")
        blk = module.blocks[0]
        cfg = DiTConfig(dim=module.dim, in_dim=module.in_dim, ffn_dim=blk.ffn_dim, out_dim=module.head.head.out_features // 4,
                        text_dim=module.text_embedding[0].in_features, freq_dim=module.freq_dim,
                        eps=blk.norm1.eps, num_heads=blk.num_heads, num_layers=len(module.blocks),
                        patch_size=tuple(module.patch_size))
        out = cls(cfg, module.state_dict(), device=device)
        out.require_vae_embedding = bool(getattr(module, "require_vae_embedding", True))
        out.require_clip_embedding = bool(getattr(module, "require_clip_embedding", False))
        return out

    # -------------------------------------------------------------------------------------------- step-invariant
    def rope(self, f: int, h: int, w: int, sp: SequenceParallel | None = None) -> torch.Tensor:
        key = (f, h, w, sp.rank if sp else 0, sp.size if sp else 1)
        t = self._rope_cache.get(key)
        if t is None:
            sl = sp.token_slice(f * h * w) if sp is not None and sp.size > 1 else None
            t = rope_cos_sin(self.cfg.head_dim, f, h, w, self.device, sl,
                             sp.rows_per_rank(f * h * w) if sl is not None else None)
            self._rope_cache = {key: t}
        return t

    def context_entry(self, context: torch.Tensor, cache: bool = True) -> tuple:
        """text_embedding (wan_video_dit.py:309-313,371) + per-block cross-attention K/V, cached per context tensor
        (the pipeline passes the same prompt embedding at every step).  Returns the cache entry
        (context, [K|V per block], embedding); the entry keeps `context` alive, so its address cannot be recycled
        while the entry is cached.  Tensors without a version counter (torch.inference_mode) are not cached."""
        try:
            key = (context.data_ptr(), context._version, tuple(context.shape), context.dtype) if cache else None
        except RuntimeError:
            key = None
        hit = self._ctx_cache.get(key) if key is not None else None
        if hit is not None:
            return hit
        ctx = context.reshape(-1, context.shape[-1]).to(device=self.device, dtype=torch.bfloat16).contiguous()
        e = capi.gemm(ctx, self.text0_w, self.text0_b, epi=capi.GF_EPI_BIAS_GELU)
        e = capi.gemm(e, self.text2_w, self.text2_b)
        entry = (context, [block_context_kv(b, self.cfg, e) for b in self.blocks], e)
        if key is not None:
            if len(self._ctx_cache) >= 4:
                self._ctx_cache.clear()
            self._ctx_cache[key] = entry
        return entry

    def context_kv(self, context: torch.Tensor) -> list:
        return self.context_entry(context)[1]

    def context_embedding(self, context: torch.Tensor) -> torch.Tensor:
        return self.context_entry(context)[2]

    def time_modulation(self, timestep: torch.Tensor):
        """t = time_embedding(sinusoidal(timestep)); t_mod = time_projection(t) (wan_video_dit.py:368-370);
        returns (block table [layers, 6, dim] = modulation + t_mod, head table [2, dim] = head.modulation + t)."""
        cfg = self.cfg
        ts = timestep.reshape(-1)[:1].to(device=self.device, dtype=torch.bfloat16).contiguous()
        e = capi.timestep_embedding(ts, cfg.freq_dim)
        t = capi.gemm(e, self.time0_w, self.time0_b, epi=capi.GF_EPI_BIAS_SILU)
        t = capi.gemm(t, self.time2_w, self.time2_b)                       # (1, dim)
        t_mod = capi.gemm(capi.silu(t), self.tproj_w, self.tproj_b)        # (1, 6*dim)
        block_tab = capi.add_rows(self.block_mod, t_mod.reshape(-1)).view(cfg.num_layers, 6, cfg.dim)
        head_tab = capi.add_rows(self.head_mod, t.reshape(-1)).view(2, cfg.dim)
        return block_tab, head_tab, t_mod.reshape(-1)

    # -------------------------------------------------------------------------------------------- pieces
    def patchify(self, x: torch.Tensor, y: torch.Tensor | None = None):
        """patch_embedding Conv3d + 'b c f h w -> b (f h w) c' for one sample: x (C0,F,H,W) [+ y (C1,F,H,W)]."""
        tok = capi.patch_gather(x, y)
        if tok.shape[1] != self.patch_w.shape[1]:
            raise ValueError(f"got {tok.shape[1] // 4} input channels, patch_embedding expects {self.cfg.in_dim}")
        F, H, W = x.shape[1:]
        return capi.gemm(tok, self.patch_w, self.patch_b), (F, H // 2, W // 2)

    def head_tokens(self, x: torch.Tensor, head_tab: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
        """Head.forward (wan_video_dit.py:262-269) on a (local) token range -> [rows, 4*out_dim]."""
        cfg = self.cfg
        ws = _workspace(x.shape[0], cfg, self.device)
        hn = capi.layernorm(x, eps=cfg.eps, shift=head_tab[0], scale=head_tab[1], out=ws.h[:x.shape[0]])
        return capi.gemm(hn, self.head_w, self.head_b, out=out)

    def unpatchify(self, tok: torch.Tensor, grid) -> torch.Tensor:
        """WanModel.unpatchify (wan_video_dit.py:351-356) -> (out_dim, F, 2h, 2w)."""
        f, h, w = grid
        return capi.unpatchify(tok, self.cfg.out_dim, f, 2 * h, 2 * w)

    # -------------------------------------------------------------------------------------------- forward
    def forward(self, x, timestep, context, clip_feature=None, y=None, use_gradient_checkpointing=False,
                use_gradient_checkpointing_offload=False, **kwargs):
        """WanModel.forward signature (wan_video_dit.py:358-367). For A14B the reference ignores clip_feature and y
        here (has_image_input=False) and expects x to carry all in_dim channels; like model_fn_wan_video (:1457) we
        additionally concatenate y when x alone does not have in_dim channels."""
        if use_gradient_checkpointing or use_gradient_checkpointing_offload:
            raise NotImplementedError("goal_force_b200 is an inference path; gradient checkpointing is not supported")
        yy = y if (y is not None and x.shape[1] != self.cfg.in_dim) else None
        return model_fn_wan_video(dit=self, latents=x, timestep=timestep, context=context, y=yy)

    __call__ = forward

    def parameters(self):
        for b in self.blocks:
            for n in _Block.__slots__:
                yield getattr(b, n)


class _PrefixTolerant:
    """state-dict view that accepts keys with or without a checkpoint prefix (load_controlnet_weights strips
    'pipe.controlnet.', src/goal_force/wan_video_new.py:176-178)."""

    def __init__(self, sd, prefix: str):
        self.sd, self.prefix = sd, prefix

    def __getitem__(self, key):
        try:
            return self.sd[key]
        except KeyError:
            return self.sd[self.prefix + key]


class ControlNetB200:
    """goal-force ControlNet (src/goal_force/wan_video_new.py:97-117) in kernel layout.
    state-dict keys: controlnet_patch_embedding.patch_embedding.*, controlnet_dit.blocks.N.*,
    controlnet_zero_convs_after.N.{weight (dim,dim,1), bias}; a leading 'pipe.controlnet.' is stripped as in
    load_controlnet_weights (:176-178)."""

    def __init__(self, cfg: DiTConfig, state_dict: dict, num_layers: int, stride=None, device="cuda"):
        _check_cfg(cfg)
        self.cfg, self.num_layers, self.stride = cfg, num_layers, stride
        self.device = torch.device(device)
        sd = _PrefixTolerant(state_dict, "pipe.controlnet.")

        def g(name):
            return sd[name].detach().to(device=self.device, dtype=torch.bfloat16).contiguous()

        self.patch_w = g("controlnet_patch_embedding.patch_embedding.weight").reshape(cfg.dim, -1).contiguous()
        self.patch_b = g("controlnet_patch_embedding.patch_embedding.bias")
        self.blocks = [_Block.from_state_dict(sd, f"controlnet_dit.blocks.{i}.", self.device) for i in range(num_layers)]
        self.block_mod = torch.stack([b.modulation for b in self.blocks], 0).contiguous() if num_layers else None
        self.zero_w = [g(f"controlnet_zero_convs_after.{i}.weight").reshape(cfg.dim, cfg.dim).contiguous()
                       for i in range(num_layers)]
        self.zero_b = [g(f"controlnet_zero_convs_after.{i}.bias") for i in range(num_layers)]
        # F6: an untrained / never-loaded ControlNet has all-zero zero-convs; its branch is then an exact no-op
        self.is_noop = stride is None and all(bool((w == 0).all()) and bool((b == 0).all())
                                              for w, b in zip(self.zero_w, self.zero_b))
        self._patch_cache: dict = {}
        self._ctx_cache: dict = {}

    @classmethod
    def from_reference(cls, module, device="cuda") -> "ControlNetB200":
        blk = module.controlnet_dit.blocks[0]
        cfg = DiTConfig(dim=blk.dim, in_dim=16, ffn_dim=blk.ffn_dim, out_dim=16,
                        text_dim=4096, freq_dim=256, eps=blk.norm1.eps, num_heads=blk.num_heads,
                        num_layers=module.num_layers)
        return cls(cfg, module.state_dict(), module.num_layers, stride=module.stride, device=device)

    def control_tokens(self, control_latents: torch.Tensor, cache: bool = True) -> torch.Tensor:
        """ControlNet_PatchEmbedding (wan_video_new.py:72-94); step-invariant, cached per latent tensor."""
        try:
            key = (control_latents.data_ptr(), control_latents._version, tuple(control_latents.shape)) if cache else None
        except RuntimeError:
            key = None
        hit = self._patch_cache.get(key) if key is not None else None   # entry keeps the latent tensor alive
        if hit is not None:
            return hit[1]
        c = control_latents
        if c.dim() == 5:
            if c.shape[0] != 1:
                raise NotImplementedError("control_signal_video_latents batch > 1")
            c = c[0]
        c = c.to(device=self.device, dtype=torch.bfloat16).contiguous()
        tok = capi.gemm(capi.patch_gather(c, None), self.patch_w, self.patch_b)
        if key is not None:
            self._patch_cache = {key: (control_latents, tok)}
        return tok

    def context_kv(self, entry: tuple, cache: bool = True) -> list:
        """Cross-attention K|V of the ControlNet blocks for one prompt.  `entry` is the trunk's cache entry
        (WanModelB200.context_entry): the ControlNet result is keyed by that object's identity and keeps it -- and
        with it the context tensor -- alive, so a recycled address or a trunk-side eviction can never alias a stale
        ControlNet entry."""
        hit = self._ctx_cache.get(id(entry)) if cache else None
        if hit is not None and hit[0] is entry:
            return hit[1]
        kvs = [block_context_kv(b, self.cfg, entry[2]) for b in self.blocks]
        if not cache:
            return kvs
        if len(self._ctx_cache) >= 4:
            self._ctx_cache.clear()
        self._ctx_cache[id(entry)] = (entry, kvs)
        return kvs


_SIDE_STREAMS: dict = {}


def _side_stream(device) -> "torch.cuda.Stream":
    key = str(device)
    st = _SIDE_STREAMS.get(key)
    if st is None:
        st = torch.cuda.Stream(device=device)
        _SIDE_STREAMS[key] = st
    return st


_CONVERTED: dict = {}      # id(reference module) -> (weakref, fingerprint, converted)
_DEFAULT_SP: list = []


def _default_sequence_parallel() -> "SequenceParallel | None":
    """`use_unified_sequence_parallel=True` without an explicit SequenceParallel object (what the reference pipeline
    passes after enable_usp, diffsynth/pipelines/wan_video_new.py:298-310): shard over the default process group."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return None
    if not _DEFAULT_SP:
        _DEFAULT_SP.append(SequenceParallel(None, transport="peer"))
    return _DEFAULT_SP[0]


def _fingerprint(module) -> tuple:
    """Cheap identity of a reference module's weights: storage address and in-place version counter of every
    parameter / buffer.  load_state_dict, LoRA merges, .to(device) and optimiser steps all change it."""
    return tuple((t.data_ptr(), t._version, t.device.index) for t in module.state_dict(keep_vars=True).values())


def _module_device(module):
    """Device the converted copy lives on: the module's own CUDA device, else the current CUDA device."""
    for t in module.parameters():
        if t.is_cuda:
            return t.device
        break
    return torch.device("cuda", torch.cuda.current_device())


def convert(obj, kind=None, device=None):
    """Explicit conversion / refresh of a reference nn.Module (WanModel or the goal-force ControlNet) into kernel
    layout.  model_fn_wan_video calls this lazily; call it yourself after changing weights in a way the fingerprint
    cannot see, or to convert ahead of the first step.  The cache holds the reference module only weakly."""
    import weakref
    if kind is None:
        kind = ControlNetB200 if hasattr(obj, "controlnet_dit") else WanModelB200
    key = id(obj)
    conv = kind.from_reference(obj, device=device or _module_device(obj))

    def _drop(_ref, key=key):
        _CONVERTED.pop(key, None)

    _CONVERTED[key] = (weakref.ref(obj, _drop), _fingerprint(obj), conv)
    return conv


def _as_b200(obj, kind):
    """Accept our classes, or reference nn.Modules.  A module is converted on first use and re-converted whenever
    its weights change (pipe.load_lora, load_controlnet_weights after a warm-up call, .to(device)): the cache entry is
    keyed by identity, validated by a weight fingerprint, and `ControlNetB200.is_noop` is recomputed with it."""
    if obj is None or isinstance(obj, kind):
        return obj
    hit = _CONVERTED.get(id(obj))
    if hit is not None and hit[0]() is obj and hit[1] == _fingerprint(obj):
        return hit[2]
    return convert(obj, kind)


def model_fn_wan_video(dit=None, motion_controller=None, vace=None, latents=None, timestep=None, context=None,
                       clip_feature=None, y=None, reference_latents=None, vace_context=None, vace_scale=1.0,
                       audio_embeds=None, motion_latents=None, s2v_pose_latents=None, drop_motion_frames=True,
                       tea_cache=None, use_unified_sequence_parallel=False, motion_bucket_id=None,
                       sliding_window_size=None, sliding_window_stride=None, cfg_merge=False,
                       use_gradient_checkpointing=False, use_gradient_checkpointing_offload=False,
                       control_camera_latents_input=None, fuse_vae_embedding_in_latents=False, controlnet=None,
                       sequence_parallel: SequenceParallel | None = None, controlnet_stream: bool | None = None,
                       _recompute_step_invariants: bool = False, **kwargs):
    """Drop-in for pipe.model_fn (src/goal_force/wan_video_new.py:1349-1591), goal-force configuration.

    Same keyword surface as the reference; extra keywords are accepted and ignored (the pipeline passes its whole
    shared-input dict). Options that select other Wan variants raise NotImplementedError instead of silently
    computing something else. Returns a new (B, out_dim, F, H, W) bf16 tensor; inputs are not modified.
    `dit` / `controlnet` may be WanModelB200 / ControlNetB200 or the reference nn.Modules (converted on first use).
    `controlnet_stream`: run the ControlNet branch on a second CUDA stream next to the trunk (the branch
    never reads the noisy latents; the trunk only needs state i after its own block i).  None = on under sequence
    parallelism, where one branch's exchange waits are filled by the other branch's kernels, off on a single GPU
    (every kernel fills the GPU there, the two streams would only interleave).  Results are bit-identical either way.
    """
    for name, val in (("motion_controller", motion_controller if motion_bucket_id is not None else None),
                      ("vace_context", vace_context), ("audio_embeds", audio_embeds),
                      ("reference_latents", reference_latents), ("tea_cache", tea_cache),
                      ("sliding_window_size", sliding_window_size),
                      ("control_camera_latents_input", control_camera_latents_input)):
        if val is not None:
            raise NotImplementedError(f"{name} is outside the goal-force hot path implemented by goal_force_b200")
    if use_gradient_checkpointing or use_gradient_checkpointing_offload:
        raise NotImplementedError("inference only")
    if latents is None or timestep is None or context is None:
        raise ValueError("latents, timestep and context are required")
    dit = _as_b200(dit, WanModelB200)
    controlnet = _as_b200(controlnet, ControlNetB200)
    if clip_feature is not None and dit.require_clip_embedding:
        raise NotImplementedError("clip_feature branch (has_image_input) is not part of the A14B hot path")
    cfg = dit.cfg
    if sequence_parallel is None and use_unified_sequence_parallel:
        sequence_parallel = _default_sequence_parallel()
    sp = sequence_parallel if (sequence_parallel is not None and sequence_parallel.size > 1) else None
    if sp is not None:
        sp.check()                      # a barrier of an earlier forward timed out -> raise instead of going on

    B = max(latents.shape[0], context.shape[0])            # merged-CFG batches replicate latents (:1451-1454)
    outs = []
    block_tab, head_tab, t_mod_flat = dit.time_modulation(timestep)
    for b in range(B):
        lat = latents[min(b, latents.shape[0] - 1)].to(device=dit.device, dtype=torch.bfloat16).contiguous()
        yb = None
        if y is not None and dit.require_vae_embedding:
            yb = y[min(b, y.shape[0] - 1)].to(device=dit.device, dtype=torch.bfloat16).contiguous()
        ctx_b = context[b:b + 1] if context.shape[0] > 1 else context
        use_cache = not _recompute_step_invariants      # graph capture: every tensor input is a static buffer
        ctx_entry = dit.context_entry(ctx_b, cache=use_cache)
        ctx_kv = ctx_entry[1]
        x, (f, h, w) = dit.patchify(lat, yb)                                                    # :1464
        L = f * h * w
        kv_len = None
        if sp is not None:
            x = sp.shard_rows(x, L)                                                             # :1526-1531
            if L % sp.size:
                kv_len = L                       # padded tail rows exist on the last rank(s): never attend to them
        cos_sin = dit.rope(f, h, w, sp)
        ws = _workspace(x.shape[0], cfg, dit.device)

        states, side, cn_events = None, None, []
        use_cn = controlnet is not None and not controlnet.is_noop
        if use_cn:                                                                              # :1489-1522
            csl = kwargs.get("control_signal_video_latents", None)
            if csl is None:
                raise ValueError("controlnet given but control_signal_video_latents is missing")
            s = controlnet.control_tokens(csl, cache=use_cache)
            if s.shape[0] != L:
                raise ValueError("control latents and latents have different token counts")
            if sp is not None:
                s = sp.shard_rows(s, L)
            cn_kv = controlnet.context_kv(ctx_entry, cache=use_cache)
            # ControlNet blocks share t_mod with the trunk but carry their own modulation tables
            cn_tab = capi.add_rows(controlnet.block_mod, t_mod_flat).view(controlnet.num_layers, 6, cfg.dim)
            states = ws.cn_states(controlnet.num_layers)       # block i: state i-1 (or the patch tokens) -> state i
            side = None
            if controlnet_stream if controlnet_stream is not None else sp is not None:
                side = _side_stream(dit.device)
            if side is None:
                for i, bw in enumerate(controlnet.blocks):
                    with capi.nvtx_range(f"controlnet.block{i}"):
                        run_block(bw, cfg, states[i], cn_kv[i], cn_tab[i], cos_sin, ws, sp, x_in=s, kv_len=kv_len)
                    s = states[i]
            else:
                # second stream, second workspace and (under sequence parallelism) second set of exchange buffers;
                # one event per ControlNet block tells the trunk that state i is complete
                main = torch.cuda.current_stream()
                ws_cn = _workspace(x.shape[0], cfg, dit.device, slot=1)
                side.wait_stream(main)
                cn_events = []
                with torch.cuda.stream(side):
                    for i, bw in enumerate(controlnet.blocks):
                        with capi.nvtx_range(f"controlnet.block{i}"):
                            run_block(bw, cfg, states[i], cn_kv[i], cn_tab[i], cos_sin, ws_cn, sp, x_in=s, slot=1,
                                      kv_len=kv_len)
                        s = states[i]
                        ev = torch.cuda.Event()
                        ev.record(side)
                        cn_events.append(ev)

        for i, bw in enumerate(dit.blocks):                                                     # :1540-1570
            with capi.nvtx_range(f"trunk.block{i}"):
                run_block(bw, cfg, x, ctx_kv[i], block_tab[i], cos_sin, ws, sp, kv_len=kv_len)
            if use_cn:
                if side is not None:
                    j = i if controlnet.stride is None else (i // controlnet.stride if i % controlnet.stride == 0 else -1)
                    if 0 <= j < len(cn_events):
                        torch.cuda.current_stream().wait_event(cn_events[j])
                if controlnet.stride is not None:
                    if i % controlnet.stride == 0 and i // controlnet.stride < len(states):
                        capi.add_(x, states[i // controlnet.stride])
                elif i < controlnet.num_layers:
                    capi.gemm(states[i], controlnet.zero_w[i], controlnet.zero_b[i], epi=capi.GF_EPI_GATE_RES,
                              gate=None, residual=x, out=x)
        if use_cn and side is not None:
            torch.cuda.current_stream().wait_stream(side)      # the next call may reuse the branch's buffers
        tok = dit.head_tokens(x, head_tab)                                                      # :1581
        if sp is not None:                                                                      # :1582-1585
            full = torch.empty((sp.rows_per_rank(L) * sp.size, tok.shape[1]), dtype=torch.bfloat16, device=dit.device)
            sp.dist.all_gather_into_tensor(full, tok, group=sp.group)
            tok = full[:L]                                        # strips the padding rows (reference :1584-1585)
        outs.append(dit.unpatchify(tok, (f, h, w)))                                             # :1590
    return torch.stack(outs, 0)
