"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: Ulysses head/sequence exchange and the CFG axis.

The product code has no CPU path, so the three kernels SequenceParallel calls (ulysses_pack, attention,
ulysses_unpack) and cfg_euler are replaced INSIDE THE TEST PROCESSES by torch restatements of their documented
layouts (include/goalforce_b200.h); what is under test is everything around them: token slicing, buffer shapes and
pitches, all-to-all send/receive ordering, group construction, and the CFG noise-prediction exchange.
"""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _attn_ref(q, k, v, heads):
    import torch.nn.functional as F
    Lq = q.shape[0]
    qh, kh, vh = (t.float().reshape(t.shape[0], heads, 128).transpose(0, 1)[None] for t in (q, k, v))
    return F.scaled_dot_product_attention(qh, kh, vh)[0].transpose(0, 1).reshape(Lq, heads * 128).to(q.dtype)


def _install_kernel_doubles():
    """torch restatements of the C-ABI layouts, installed over goal_force_b200.capi in this process only."""
    from goal_force_b200 import capi

    def ulysses_pack(x, heads, head_dim, P, out=None, out_pitch=None):
        rows = x.shape[0]
        w = (heads // P) * head_dim
        out_pitch = w if out_pitch is None else out_pitch
        if out is None:
            out = torch.empty((P, rows, out_pitch), dtype=x.dtype)
        # out may be a column-offset view of a wider [P, rows, pitch] send buffer: write through strides
        src = x[:, :heads * head_dim].reshape(rows, P, w).permute(1, 0, 2)
        out[:, :, :w] = src
        return out

    def ulysses_unpack(inp, rows, heads, head_dim, P, out=None):
        w = (heads // P) * head_dim
        res = inp.reshape(P, rows, w).permute(1, 0, 2).reshape(rows, heads * head_dim)
        if out is None:
            return res.contiguous()
        out.copy_(res)
        return out

    def attention(q, k, v, heads, out=None, scale=None, kv_len=None):
        if kv_len is not None:                       # zero-padded token tail: attend to the real keys only
            k, v = k[:kv_len], v[:kv_len]
        res = _attn_ref(q, k, v, heads)
        if out is None:
            return res
        out.copy_(res)
        return out

    def cfg_euler(posi, nega, latents, cfg_scale, dsigma, out=None):
        pred = posi if nega is None else nega + cfg_scale * (posi - nega)
        res = latents + pred * torch.tensor(dsigma)
        if out is not None:
            out.copy_(res)
            return out
        return res

    capi.ulysses_pack, capi.ulysses_unpack, capi.attention, capi.cfg_euler = (ulysses_pack, ulysses_unpack, attention,
                                                                              cfg_euler)


def _init(rank, world, port):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    return dist


def _worker_ulysses(rank, world, port, L, heads):
    dist = _init(rank, world, port)
    try:
        _install_kernel_doubles()
        from goal_force_b200.wan_dit import SequenceParallel
        sp = SequenceParallel(None)
        d = heads * 128
        g = torch.Generator().manual_seed(0)
        qkv = torch.randn(L, 3 * d, generator=g).bfloat16()            # identical on every rank
        full = _attn_ref(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], heads)
        sl = sp.token_slice(L)
        assert (sl.stop - sl.start) == L // world and sl.start == rank * (L // world)
        out = torch.empty(L // world, d, dtype=torch.bfloat16)
        sp.self_attention(qkv[sl].contiguous(), heads, out)
        assert torch.equal(out, full[sl])
        # second call reuses the cached exchange buffers
        sp.self_attention(qkv[sl].contiguous(), heads, out)
        assert torch.equal(out, full[sl])
        # a token count that does not divide: ceil(L/P) rows per rank, zero-padded tail, padded keys masked (kv_len)
        Lo = L + 1
        n = sp.rows_per_rank(Lo)
        assert n == -(-Lo // world)
        qkv_o = torch.randn(Lo, 3 * d, generator=g).bfloat16()
        full_o = _attn_ref(qkv_o[:, :d], qkv_o[:, d:2 * d], qkv_o[:, 2 * d:], heads)
        slo = sp.token_slice(Lo)
        assert slo.start == min(rank * n, Lo) and slo.stop == min((rank + 1) * n, Lo)
        out_o = torch.empty(n, d, dtype=torch.bfloat16)
        sp.self_attention(sp.shard_rows(qkv_o, Lo), heads, out_o, 0, Lo)
        assert torch.equal(out_o[: slo.stop - slo.start], full_o[slo])
        with pytest.raises(ValueError):
            sp.self_attention(torch.randn(4, 3 * 3 * 128).bfloat16(), 3, torch.empty(4, 3 * 128).bfloat16())   # 3 heads over 2 ranks
    finally:
        dist.destroy_process_group()


def _worker_cfg(rank, world, port):
    dist = _init(rank, world, port)
    try:
        _install_kernel_doubles()
        from goal_force_b200.pipeline import GoalForceDenoiser, ParallelContext, ParallelLayout
        from goal_force_b200.scheduler import FlowMatchScheduler
        par = ParallelContext(ParallelLayout(world_size=world, rank=rank, cfg_size=2))
        assert par.sp is None and par.layout.cfg_index == rank
        calls = []

        def fake_model_fn(dit=None, latents=None, timestep=None, context=None, **kw):
            calls.append(float(context.flatten()[0]))
            return latents * context.flatten()[0]                     # "prediction" identifies the branch

        class _Dit:  # only identity matters to the denoiser
            pass

        den = GoalForceDenoiser(_Dit(), parallel=par, model_fn=fake_model_fn,
                                scheduler=FlowMatchScheduler(shift=5, sigma_min=0.0, extra_one_step=True))
        den.scheduler.set_timesteps(4, shift=5.0)
        lat = torch.arange(24, dtype=torch.float32).reshape(1, 2, 3, 4)
        posi_ctx, nega_ctx = torch.full((1, 1, 1), 2.0), torch.full((1, 1, 1), -1.0)
        t = den.scheduler.timesteps[1]
        got = den.step(lat, t, posi_ctx, nega_ctx, None, None, cfg_scale=5.0)
        # each rank ran exactly one branch; the combined result equals the sequential formula on every rank
        assert calls == [2.0 if rank == 0 else -1.0]
        p, n = lat * 2.0, lat * -1.0
        want = lat + (n + 5.0 * (p - n)) * torch.tensor(den.scheduler.dsigma(t))
        torch.testing.assert_close(got, want)
        # cfg_scale == 1: no exchange, only the conditional branch, on every rank
        calls.clear()
        got = den.step(lat, t, posi_ctx, nega_ctx, None, None, cfg_scale=1.0)
        assert calls == [2.0]
        torch.testing.assert_close(got, lat + p * torch.tensor(den.scheduler.dsigma(t)))
    finally:
        dist.destroy_process_group()


def _spawn(fn, *args):
    port = _free_port()
    mp.spawn(fn, args=(2, port, *args), nprocs=2, join=True)


@pytest.mark.timeout(300)
def test_ulysses_sequence_parallel_equals_single_rank():
    _spawn(_worker_ulysses, 64, 4)


@pytest.mark.timeout(300)
def test_ulysses_ragged_head_groups():
    _spawn(_worker_ulysses, 48, 2)          # one head per rank


@pytest.mark.timeout(300)
def test_cfg_parallel_equals_sequential():
    _spawn(_worker_cfg)
