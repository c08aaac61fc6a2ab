"""CUDA-graph capture of the DiT forward (`gf_ctx` + graph capture of one forward per (L, expert)).

One A14B forward is ~770 launches through ctypes; replaying a captured graph removes the Python / ctypes / launch
overhead from the step entirely (it matters at 8 GPUs, where a forward is ~240 ms, and for short sequences).

`GraphedModelFn` is a drop-in for `model_fn_wan_video` (it can be handed to GoalForceDenoiser(model_fn=...) or
assigned to `pipe.model_fn`).  One graph is captured per (expert, ControlNet, tensor shapes, sequence-parallel group);
ALL tensor inputs -- latents, timestep, context, y, control latents -- are copied into static buffers before a replay,
and the step-invariant pieces (text embedding, per-block cross-attention K/V, ControlNet patch tokens) are recomputed
inside the graph (0.15 % of the forward's FLOPs), so a graph stays valid for any prompt, image and control signal of
the same shape.  Everything a capture must not contain -- the RoPE table upload, cudaFuncSetAttribute, the peer-memory
exchange set-up, TMA descriptor encoding misses -- happens in the eager warm-up call that precedes the capture.
"""
from __future__ import annotations

import torch

from . import capi
from .wan_dit import model_fn_wan_video

_TENSOR_KEYS = ("latents", "timestep", "context", "y", "control_signal_video_latents")


class GraphedModelFn:
    def __init__(self, model_fn=model_fn_wan_video, warmup: int = 1):
        self.model_fn = model_fn
        self.warmup = warmup
        self._graphs: dict = {}
        self.captures = 0
        self.replays = 0

    @staticmethod
    def _key(kw: dict):
        parts = [id(kw.get("dit")), id(kw.get("controlnet")), id(kw.get("sequence_parallel")),
                 bool(kw.get("use_unified_sequence_parallel", False))]
        for k in _TENSOR_KEYS:
            t = kw.get(k)
            parts.append(None if t is None else (tuple(t.shape), t.dtype, str(t.device)))
        return tuple(parts)

    def __call__(self, **kw) -> torch.Tensor:
        if capi.STATS.timing:
            raise RuntimeError("per-launch event timing (capi.STATS.timing) cannot be combined with graph replay")
        key = self._key(kw)
        entry = self._graphs.get(key)
        if entry is None:
            entry = self._capture(kw)
            self._graphs[key] = entry
        graph, static, out = entry
        for k, buf in static.items():
            buf.copy_(kw[k], non_blocking=True)
        graph.replay()
        self.replays += 1
        return out.clone()

    def _capture(self, kw: dict):
        static = {k: kw[k].clone() for k in _TENSOR_KEYS if kw.get(k) is not None}
        call = dict(kw)
        call.update(static)
        call["_recompute_step_invariants"] = True
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(max(1, self.warmup)):          # eager: fills host caches, workspaces, exchange buffers
                self.model_fn(**call)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            out = self.model_fn(**call)
        self.captures += 1
        return graph, static, out
