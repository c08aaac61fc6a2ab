"""ctypes binding of libgoalforce_b200.so (the C ABI in include/goalforce_b200.h).

torch is used only to own device memory and streams: every call hands raw device pointers and the current
CUDA stream to the library. There is no CPU or PyTorch fallback: if the library is missing or a call fails the
wrappers raise.
"""
from __future__ import annotations

import ctypes
import os
from pathlib import Path

import torch

_LIB = None

GF_EPI_BIAS = 0
GF_EPI_BIAS_GELU = 1
GF_EPI_BIAS_SILU = 2
GF_EPI_GATE_RES = 3
GF_EPI_F32 = 4

ABI_VERSION = 3

_ERRORS = {-1: "GF_ERR_BAD_ARG", -2: "GF_ERR_NO_DRIVER", -3: "GF_ERR_TMAP", -4: "GF_ERR_UNSUPPORTED"}

# name -> argtypes; every entry point returns int
_p, _ll, _i, _f = ctypes.c_void_p, ctypes.c_longlong, ctypes.c_int, ctypes.c_float
SIGNATURES = {
    "gf_abi_version": [],
    "gf_device_sms": [],
    "gf_ctx_create": [ctypes.POINTER(ctypes.c_void_p)],
    "gf_ctx_destroy": [_p],
    "gf_ctx_set_attention": [_p, _i, _i],
    "gf_ctx_set_gemm_raster": [_p, _i],
    "gf_ctx_set_gemm_tile": [_p, _i],
    "gf_ctx_set_conv": [_p, _i],
    "gf_ctx_stats": [_p, ctypes.POINTER(_ll), ctypes.POINTER(_ll), ctypes.POINTER(_ll)],
    "gf_gemm_bf16": [_p, _p, _ll, _p, _ll, _p, _ll, _i, _i, _i, _p, _i, _p, _p, _ll, _i, _p],
    "gf_layernorm_bf16": [_p, _ll, _p, _ll, _i, _i, _f, _p, _p, _p, _p, _p],
    "gf_rmsnorm_rope_bf16": [_p, _ll, _i, _i, _p, _f, _p, _i, _p],
    "gf_qk_rmsnorm_rope_bf16": [_p, _ll, _i, _i, _p, _p, _f, _p, _i, _p],
    "gf_attention_bf16": [_p, _p, _ll, _p, _ll, _p, _ll, _p, _ll, _i, _i, _i, _i, _f, _p],
    "gf_patch_gather_bf16": [_p, _i, _p, _i, _p, _ll, _i, _i, _i, _p],
    "gf_unpatchify_bf16": [_p, _ll, _p, _i, _i, _i, _i, _p],
    "gf_add_rows_bf16": [_p, _p, _p, _i, _i, _p],
    "gf_add_bf16": [_p, _p, _p, _ll, _p],
    "gf_silu_bf16": [_p, _p, _ll, _p],
    "gf_cfg_euler_bf16": [_p, _p, _p, _p, _f, _f, _ll, _p],
    "gf_timestep_embedding_bf16": [_p, _p, _i, _i, _p],
    "gf_embedding_bf16": [_p, _p, _p, _i, _i, _ll, _p],
    "gf_t5_rmsnorm_bf16": [_p, _ll, _p, _ll, _i, _i, _p, _f, _p],
    "gf_mul_bf16": [_p, _p, _p, _ll, _p],
    "gf_t5_attention_bf16": [_p, _ll, _p, _ll, _p, _ll, _p, _ll, _i, _i, _i, _i, _i, _p, _p, _p, _p],
    "gf_conv3d_cl_bf16": [_p, _p, _ll, _i, _i, _i, _i, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p, _p, _ll, _i, _i, _i,
                          _p, _ll, _p, _ll, _p, _i, _i, _p],
    "gf_vae_rmsnorm_bf16": [_p, _ll, _p, _ll, _ll, _i, _p, _i, _p],
    "gf_vae_upsample2x_bf16": [_p, _ll, _p, _ll, _p, _ll, _i, _i, _i, _i, _p],
    "gf_softmax_f32_bf16": [_p, _ll, _p, _ll, _i, _i, _i, _f, _p],
    "gf_vae_planes_to_cl_bf16": [_p, _ll, _i, _p, _ll, _i, _p, _p, _i, _i, _i, _p],
    "gf_vae_cl_to_planes_bf16": [_p, _ll, _ll, _i, _p, _p, _p, _i, _p],
    "gf_vae_head_gather_bf16": [_p, _ll, _p, _p, _i, _i, _i, _i, _p],
    "gf_vae_blend_bf16": [_p, _i, _i, _i, _i, _p, _i, _i, _i, _i, _p, _p],
    "gf_vae_blend_finish_bf16": [_p, _ll, _i, _i, _p, _i, _p],
    "gf_peer_alloc": [ctypes.POINTER(ctypes.c_void_p), _ll],
    "gf_peer_free": [_p],
    "gf_peer_export": [_p, _p],
    "gf_peer_import": [_p, ctypes.POINTER(ctypes.c_void_p)],
    "gf_peer_unimport": [_p],
    "gf_peer_barrier": [ctypes.POINTER(ctypes.c_void_p), _i, _i, _ll, _p, _p],
    "gf_peer_status_alloc": [ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_void_p)],
    "gf_peer_status_free": [_p],
    "gf_qkv_rmsnorm_rope_scatter_bf16": [_p, _ll, _i, _i, _p, _p, _f, _p, _i, ctypes.POINTER(ctypes.c_void_p), _i, _i,
                                         _ll, _p],
    "gf_attention_scatter_bf16": [_p, _p, _ll, _p, _ll, _p, _ll, ctypes.POINTER(ctypes.c_void_p), _i, _ll, _i, _i, _i,
                                  _i, _i, _i, _f, _p],
    "gf_ulysses_pack_bf16": [_p, _ll, _p, _ll, _i, _i, _i, _i, _p],
    "gf_ulysses_unpack_bf16": [_p, _p, _ll, _i, _i, _i, _i, _p],
}


def lib_path() -> Path:
    """In-tree library; GF_B200_LIB points at another build of the SAME ABI (kernel A/B experiments)."""
    override = os.environ.get("GF_B200_LIB")
    return Path(override) if override else Path(__file__).resolve().parent / "_lib" / "libgoalforce_b200.so"


def load() -> ctypes.CDLL:
    """Load the shared library (building it first when GF_B200_AUTOBUILD=1 and it is absent)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not path.exists():
        if os.environ.get("GF_B200_AUTOBUILD", "0") == "1":
            from . import build as _build
            _build.build()
        else:
            raise RuntimeError(
                f"{path} is missing: run `python -m goal_force_b200.build` (or __graft_entry__.build()). "
                "goal_force_b200 has no CPU/PyTorch fallback.")
    lib = ctypes.CDLL(str(path))
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)  # raises AttributeError if a declared symbol is not exported
        fn.argtypes = argtypes
        fn.restype = ctypes.c_int
    if lib.gf_abi_version() != ABI_VERSION:
        raise RuntimeError("libgoalforce_b200.so ABI version mismatch")
    _LIB = lib
    return lib


_CTX: dict = {}


def ctx() -> int:
    """The gf_ctx of this process for the current CUDA device (descriptor cache + tuning), created on first use.
    Initial tuning comes from the environment, read HERE once and never on the launch path:
    GF_ATTN_IMPL (80 | 128), GF_ATTN_EMU_PAIRS (0 | 2 | 4 | 6), GF_GEMM_GROUP_M."""
    dev = torch.cuda.current_device()
    c = _CTX.get(dev)
    if c is None:
        out = ctypes.c_void_p()
        _check(load().gf_ctx_create(ctypes.byref(out)), "gf_ctx_create")
        c = out.value
        _CTX[dev] = c
        impl = int(os.environ.get("GF_ATTN_IMPL", "0"))
        emu = int(os.environ.get("GF_ATTN_EMU_PAIRS", "-1"))
        if impl or emu >= 0:
            _check(load().gf_ctx_set_attention(c, impl, emu), "gf_ctx_set_attention")
        gb = int(os.environ.get("GF_GEMM_BN", "0"))
        if gb:
            _check(load().gf_ctx_set_gemm_tile(c, gb), "gf_ctx_set_gemm_tile")
        ci = int(os.environ.get("GF_CONV_IMPL", "0"))
        if ci:
            _check(load().gf_ctx_set_conv(c, ci), "gf_ctx_set_conv")
        gm = int(os.environ.get("GF_GEMM_GROUP_M", "0"))
        if gm:
            _check(load().gf_ctx_set_gemm_raster(c, gm), "gf_ctx_set_gemm_raster")
    return c


def ctx_stats() -> dict:
    e, h, m = _ll(), _ll(), _ll()
    _check(load().gf_ctx_stats(ctx(), ctypes.byref(e), ctypes.byref(h), ctypes.byref(m)), "gf_ctx_stats")
    return {"tmap_entries": e.value, "tmap_hits": h.value, "tmap_misses": m.value}


def _check(rc: int, what: str) -> None:
    if rc != 0:
        name = _ERRORS.get(rc, f"cudaError {rc}")
        raise RuntimeError(f"{what} failed: {name}")


class LaunchStats:
    """Counts kernel launches made through this binding and, when `timing` is on, brackets every launch with CUDA
    events on the launching stream (used by bench.py for the per-kernel roofline numbers)."""

    def __init__(self):
        self.launches = 0
        self.timing = False
        self.records: list = []      # (tag, start_event, end_event, work) ; work = algorithmic flops or bytes

    def reset(self, timing: bool = False):
        self.launches = 0
        self.timing = timing
        self.records = []

    def summary(self) -> dict:
        """tag -> {"launches", "ms", "work"}; call after torch.cuda.synchronize()."""
        out: dict = {}
        for tag, e0, e1, work in self.records:
            d = out.setdefault(tag, {"launches": 0, "ms": 0.0, "work": 0.0})
            d["launches"] += 1
            d["ms"] += e0.elapsed_time(e1)
            d["work"] += work
        return out


STATS = LaunchStats()

# NVTX ranges (tracing): GF_B200_NVTX=1 (read once at import) or capi.NVTX = True wraps every launch in a
# range named after its kernel class and lets the host code mark forward / block / branch ranges with nvtx_range().
NVTX = os.environ.get("GF_B200_NVTX", "0") == "1"


class nvtx_range:
    """`with capi.nvtx_range("trunk.block3"):` -- a no-op unless capi.NVTX is set."""

    def __init__(self, name: str):
        self.name, self.on = name, NVTX

    def __enter__(self):
        if self.on:
            torch.cuda.nvtx.range_push(self.name)
        return self

    def __exit__(self, *exc):
        if self.on:
            torch.cuda.nvtx.range_pop()
        return False


def _call(tag: str, work: float, fn, *args) -> None:
    STATS.launches += 1
    if NVTX:
        torch.cuda.nvtx.range_push("gf." + tag)
        try:
            _check(fn(*args), tag)
        finally:
            torch.cuda.nvtx.range_pop()
        return
    if STATS.timing:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = fn(*args)
        e1.record()
        STATS.records.append((tag, e0, e1, work))
    else:
        rc = fn(*args)
    _check(rc, tag)


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: torch.Tensor | None) -> int | None:
    return None if t is None else t.data_ptr()


def _req(t: torch.Tensor, name: str, dtype=torch.bfloat16) -> None:
    if not t.is_cuda:
        raise ValueError(f"{name} must be a CUDA tensor (goal_force_b200 has no CPU path)")
    if t.dtype != dtype:
        raise ValueError(f"{name} must be {dtype}, got {t.dtype}")
    if t.dim() >= 1 and t.stride(-1) != 1:
        raise ValueError(f"{name} must be contiguous in its last dimension")


def _ld(t: torch.Tensor) -> int:
    return t.stride(0) if t.dim() == 2 else t.shape[-1]


# ------------------------------------------------------------------------------------------------ wrappers
def gemm(a: torch.Tensor, w: torch.Tensor, bias: torch.Tensor | None = None, *, epi: int = GF_EPI_BIAS,
         gate: torch.Tensor | None = None, residual: torch.Tensor | None = None,
         out: torch.Tensor | None = None, cta_group: int = 2) -> torch.Tensor:
    """out[M,N] = epi(a[M,K] @ w[N,K]^T): F.linear with the fused tails of GF_EPI_*."""
    _req(a, "a"); _req(w, "w")
    if a.dim() != 2 or w.dim() != 2 or a.shape[1] != w.shape[1]:
        raise ValueError(f"gemm shape mismatch: a {tuple(a.shape)} w {tuple(w.shape)}")
    M, K = a.shape
    N = w.shape[0]
    if out is None:
        out = torch.empty((M, N), dtype=torch.bfloat16, device=a.device)
    _req(out, "out")
    if out.shape != (M, N):
        raise ValueError("out has wrong shape")
    if bias is not None:
        _req(bias, "bias")
    if epi == GF_EPI_GATE_RES:
        if residual is None:
            raise ValueError("GF_EPI_GATE_RES needs a residual")
        _req(residual, "residual")
        if gate is not None:
            _req(gate, "gate")
    _call("gemm", 2.0 * M * N * K, load().gf_gemm_bf16, ctx(), a.data_ptr(), _ld(a), w.data_ptr(), _ld(w), out.data_ptr(),
          _ld(out), M, N, K, _ptr(bias), epi, _ptr(gate), _ptr(residual),
          _ld(residual) if residual is not None else 0, cta_group, _stream())
    return out


def layernorm(x: torch.Tensor, *, eps: float, shift: torch.Tensor | None = None, scale: torch.Tensor | None = None,
              weight: torch.Tensor | None = None, bias: torch.Tensor | None = None,
              out: torch.Tensor | None = None) -> torch.Tensor:
    _req(x, "x")
    rows, d = x.shape
    if out is None:
        out = torch.empty_like(x)
    for n, t in (("shift", shift), ("scale", scale), ("weight", weight), ("bias", bias)):
        if t is not None:
            _req(t, n)
    _call("layernorm", 4.0 * rows * d, load().gf_layernorm_bf16, x.data_ptr(), _ld(x), out.data_ptr(), _ld(out), rows,
          d, eps, _ptr(shift), _ptr(scale), _ptr(weight), _ptr(bias), _stream())
    return out


def rmsnorm_rope_(x: torch.Tensor, weight: torch.Tensor, *, eps: float, cos_sin: torch.Tensor | None,
                  head_dim: int, d: int | None = None) -> torch.Tensor:
    """In place on the first `d` columns of x[rows, ld]."""
    _req(x, "x"); _req(weight, "weight")
    rows = x.shape[0]
    d = weight.numel() if d is None else d
    if cos_sin is not None:
        _req(cos_sin, "cos_sin", torch.float32)
        if cos_sin.shape != (rows, head_dim // 2, 2) or not cos_sin.is_contiguous():
            raise ValueError(f"cos_sin must be contiguous [rows, head_dim/2, 2], got {tuple(cos_sin.shape)}")
    _call("rmsnorm_rope", 4.0 * rows * d, load().gf_rmsnorm_rope_bf16, x.data_ptr(), _ld(x), rows, d,
          weight.data_ptr(), eps, _ptr(cos_sin), head_dim, _stream())
    return x


def qk_rmsnorm_rope_(qkv: torch.Tensor, weight_q: torch.Tensor, weight_k: torch.Tensor, *, eps: float,
                     cos_sin: torch.Tensor | None, head_dim: int) -> torch.Tensor:
    """In place on the q (columns [0,d)) and k (columns [d,2d)) parts of a fused q|k|v buffer, one launch."""
    _req(qkv, "qkv"); _req(weight_q, "weight_q"); _req(weight_k, "weight_k")
    rows, d = qkv.shape[0], weight_q.numel()
    if qkv.shape[1] < 2 * d or weight_k.numel() != d:
        raise ValueError("qkv must hold at least q|k of width d each")
    if cos_sin is not None:
        _req(cos_sin, "cos_sin", torch.float32)
        if cos_sin.shape != (rows, head_dim // 2, 2) or not cos_sin.is_contiguous():
            raise ValueError(f"cos_sin must be contiguous [rows, head_dim/2, 2], got {tuple(cos_sin.shape)}")
    _call("rmsnorm_rope", 8.0 * rows * d, load().gf_qk_rmsnorm_rope_bf16, qkv.data_ptr(), _ld(qkv), rows, d,
          weight_q.data_ptr(), weight_k.data_ptr(), eps, _ptr(cos_sin), head_dim, _stream())
    return qkv


def attention_tuning(impl: int = 0, emu_pairs: int = -1) -> None:
    """Select the attention kernel of this process's context (0 = per shape, 80 = decoupled 80-row blocks,
    128 = aliased 128-row blocks) and the share of exponentials evaluated on the FMA pipe (-1 = kernel default)."""
    _check(load().gf_ctx_set_attention(ctx(), impl, emu_pairs), "gf_ctx_set_attention")


def gemm_tile_tuning(bn: int = 0) -> None:
    """Tile width of the CTA-pair GEMM for this process's context: 0 = per shape, 224 / 256 = forced."""
    _check(load().gf_ctx_set_gemm_tile(ctx(), bn), "gf_ctx_set_gemm_tile")


def conv_tuning(impl: int = 0) -> None:
    """Convolution kernel choice of this process's context: impl 0 = per shape, 1 = always tap-by-tap."""
    _check(load().gf_ctx_set_conv(ctx(), impl), "gf_ctx_set_conv")


def gemm_tuning(group_m: int = 0) -> None:
    _check(load().gf_ctx_set_gemm_raster(ctx(), group_m), "gf_ctx_set_gemm_raster")


def attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, heads: int, *, out: torch.Tensor | None = None,
              scale: float | None = None, kv_len: int | None = None) -> torch.Tensor:
    """q: [Lq, >=heads*128] view, k/v: [Lk, ...] views (row pitch taken from stride(0)).  kv_len < Lk attends only to
    the first kv_len keys (zero-padded token tails under sequence parallelism)."""
    _req(q, "q"); _req(k, "k"); _req(v, "v")
    head_dim = 128
    Lq, Lk = q.shape[0], k.shape[0]
    if kv_len is not None:
        if not 0 < kv_len <= Lk:
            raise ValueError("kv_len must be in (0, number of key rows]")
        Lk = kv_len
    if out is None:
        out = torch.empty((Lq, heads * head_dim), dtype=torch.bfloat16, device=q.device)
    if scale is None:
        scale = head_dim ** -0.5
    _call("attention_self" if Lk >= Lq else "attention_cross", 4.0 * Lq * Lk * heads * head_dim,
          load().gf_attention_bf16, ctx(), q.data_ptr(), _ld(q), k.data_ptr(), _ld(k), v.data_ptr(), _ld(v), out.data_ptr(),
          _ld(out), Lq, Lk, heads, head_dim, scale, _stream())
    return out


def patch_gather(src0: torch.Tensor, src1: torch.Tensor | None, out: torch.Tensor | None = None) -> torch.Tensor:
    """src*: (C, F, H, W) contiguous bf16 -> tokens [F*H/2*W/2, (C0+C1)*4]."""
    _req(src0, "src0")
    C0, F, H, W = src0.shape
    C1 = 0
    if src1 is not None:
        _req(src1, "src1")
        C1 = src1.shape[0]
        if tuple(src1.shape[1:]) != (F, H, W):
            raise ValueError("src1 spatial shape mismatch")
    if not src0.is_contiguous() or (src1 is not None and not src1.is_contiguous()):
        raise ValueError("patch_gather sources must be contiguous")
    L = F * (H // 2) * (W // 2)
    if out is None:
        out = torch.empty((L, (C0 + C1) * 4), dtype=torch.bfloat16, device=src0.device)
    _call("patch_gather", 4.0 * out.numel(), load().gf_patch_gather_bf16, src0.data_ptr(), C0, _ptr(src1), C1,
          out.data_ptr(), _ld(out), F, H, W, _stream())
    return out


def unpatchify(tokens: torch.Tensor, C: int, F: int, H: int, W: int, out: torch.Tensor | None = None) -> torch.Tensor:
    _req(tokens, "tokens")
    if out is None:
        out = torch.empty((C, F, H, W), dtype=torch.bfloat16, device=tokens.device)
    _call("unpatchify", 4.0 * out.numel(), load().gf_unpatchify_bf16, tokens.data_ptr(), _ld(tokens), out.data_ptr(),
          C, F, H, W, _stream())
    return out


def add_rows(a: torch.Tensor, b: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    _req(a, "a"); _req(b, "b")
    a2 = a.reshape(-1, b.numel())
    if out is None:
        out = torch.empty_like(a2)
    _call("elementwise", 4.0 * a2.numel(), load().gf_add_rows_bf16, a2.data_ptr(), b.data_ptr(), out.data_ptr(),
          a2.shape[0], a2.shape[1], _stream())
    return out.view(a.shape)


def add_(x: torch.Tensor, other: torch.Tensor) -> torch.Tensor:
    """x += other (same shape, contiguous), one bf16 rounding per element like torch's bf16 add."""
    _req(x, "x"); _req(other, "other")
    if x.shape != other.shape or not x.is_contiguous() or not other.is_contiguous():
        raise ValueError("add_ needs two contiguous tensors of the same shape")
    _call("elementwise", 6.0 * x.numel(), load().gf_add_bf16, x.data_ptr(), other.data_ptr(), x.data_ptr(), x.numel(),
          _stream())
    return x


def silu(x: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    _req(x, "x")
    if out is None:
        out = torch.empty_like(x)
    _call("elementwise", 4.0 * x.numel(), load().gf_silu_bf16, x.data_ptr(), out.data_ptr(), x.numel(), _stream())
    return out


def cfg_euler(posi: torch.Tensor, nega: torch.Tensor | None, latents: torch.Tensor, cfg_scale: float, dsigma: float,
              out: torch.Tensor | None = None) -> torch.Tensor:
    _req(posi, "posi"); _req(latents, "latents")
    if nega is not None:
        _req(nega, "nega")
    if out is None:
        out = torch.empty_like(latents)
    _call("cfg_euler", 8.0 * latents.numel(), load().gf_cfg_euler_bf16, posi.data_ptr(), _ptr(nega),
          latents.data_ptr(), out.data_ptr(), cfg_scale, dsigma, latents.numel(), _stream())
    return out


def timestep_embedding(timestep: torch.Tensor, dim: int, out: torch.Tensor | None = None) -> torch.Tensor:
    _req(timestep, "timestep")
    B = timestep.numel()
    if out is None:
        out = torch.empty((B, dim), dtype=torch.bfloat16, device=timestep.device)
    _call("elementwise", 2.0 * B * dim, load().gf_timestep_embedding_bf16, timestep.data_ptr(), out.data_ptr(), B, dim,
          _stream())
    return out


def ulysses_pack(x: torch.Tensor, heads: int, head_dim: int, P: int, out: torch.Tensor | None = None,
                 out_pitch: int | None = None) -> torch.Tensor:
    """x[rows, >=heads*head_dim] -> out[P, rows, out_pitch]; destination p gets heads [p*heads/P, (p+1)*heads/P).
    `out` may be a view into a wider send buffer (its data_ptr is the first element written)."""
    _req(x, "x")
    rows = x.shape[0]
    width = (heads // P) * head_dim
    out_pitch = width if out_pitch is None else out_pitch
    if out is None:
        out = torch.empty((P, rows, out_pitch), dtype=torch.bfloat16, device=x.device)
    _call("ulysses_pack", 4.0 * rows * heads * head_dim, load().gf_ulysses_pack_bf16, x.data_ptr(), _ld(x),
          out.data_ptr(), out_pitch, rows, heads, head_dim, P, _stream())
    return out


def ulysses_unpack(inp: torch.Tensor, rows: int, heads: int, head_dim: int, P: int,
                   out: torch.Tensor | None = None) -> torch.Tensor:
    """inp[P, rows, heads/P, head_dim] -> out[rows, heads*head_dim]."""
    _req(inp, "inp")
    if out is None:
        out = torch.empty((rows, heads * head_dim), dtype=torch.bfloat16, device=inp.device)
    _call("ulysses_pack", 4.0 * rows * heads * head_dim, load().gf_ulysses_unpack_bf16, inp.data_ptr(), out.data_ptr(),
          _ld(out), rows, heads, head_dim, P, _stream())
    return out


# ------------------------------------------------------------------------------------------------ umT5 encoder pieces
def embedding(ids: torch.Tensor, table: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    """ids: int64 [rows] on the device; table: [vocab, dim] bf16 -> [rows, dim]."""
    _req(ids, "ids", torch.int64); _req(table, "table")
    rows, dim = ids.numel(), table.shape[1]
    if not ids.is_contiguous() or not table.is_contiguous():
        raise ValueError("embedding needs contiguous ids and table")
    if out is None:
        out = torch.empty((rows, dim), dtype=torch.bfloat16, device=table.device)
    _call("elementwise", 4.0 * rows * dim, load().gf_embedding_bf16, ids.data_ptr(), table.data_ptr(), out.data_ptr(),
          rows, dim, table.shape[0], _stream())
    return out


def t5_rmsnorm(x: torch.Tensor, weight: torch.Tensor, *, eps: float, out: torch.Tensor | None = None) -> torch.Tensor:
    _req(x, "x"); _req(weight, "weight")
    rows, d = x.shape
    if out is None:
        out = torch.empty((rows, d), dtype=torch.bfloat16, device=x.device)
    _call("rmsnorm_rope", 4.0 * rows * d, load().gf_t5_rmsnorm_bf16, x.data_ptr(), _ld(x), out.data_ptr(), _ld(out),
          rows, d, weight.data_ptr(), eps, _stream())
    return out


def mul_(x: torch.Tensor, other: torch.Tensor) -> torch.Tensor:
    """x *= other (same shape, contiguous), one bf16 rounding per element."""
    _req(x, "x"); _req(other, "other")
    if x.shape != other.shape or not x.is_contiguous() or not other.is_contiguous():
        raise ValueError("mul_ needs two contiguous tensors of the same shape")
    _call("elementwise", 6.0 * x.numel(), load().gf_mul_bf16, x.data_ptr(), other.data_ptr(), x.data_ptr(), x.numel(),
          _stream())
    return x


def t5_attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, *, batch: int, heads: int,
                 bias_table: torch.Tensor, bucket_of: torch.Tensor, key_mask: torch.Tensor | None = None,
                 out: torch.Tensor | None = None) -> torch.Tensor:
    """q: [batch*Lq, >= heads*64] view, k/v: [batch*Lk, ...]; bias_table [num_buckets, heads] bf16;
    bucket_of int32 [Lq+Lk-1]; key_mask int32 [batch, Lk] or None."""
    _req(q, "q"); _req(k, "k"); _req(v, "v"); _req(bias_table, "bias_table"); _req(bucket_of, "bucket_of", torch.int32)
    Lq, Lk = q.shape[0] // batch, k.shape[0] // batch
    if bucket_of.numel() != Lq + Lk - 1 or bias_table.shape[1] != heads or not bias_table.is_contiguous():
        raise ValueError("bucket_of must have Lq+Lk-1 entries and bias_table must be contiguous [num_buckets, heads]")
    if key_mask is not None:
        _req(key_mask, "key_mask", torch.int32)
        if key_mask.shape != (batch, Lk) or not key_mask.is_contiguous():
            raise ValueError("key_mask must be contiguous [batch, Lk]")
    if out is None:
        out = torch.empty((batch * Lq, heads * 64), dtype=torch.bfloat16, device=q.device)
    _call("attention_t5", 4.0 * batch * Lq * Lk * heads * 64, load().gf_t5_attention_bf16, q.data_ptr(), _ld(q),
          k.data_ptr(), _ld(k), v.data_ptr(), _ld(v), out.data_ptr(), _ld(out), batch, Lq, Lk, heads, 64,
          bias_table.data_ptr(), bucket_of.data_ptr(), _ptr(key_mask), _stream())
    return out


# ------------------------------------------------------------------------------------------------ Wan VAE pieces
def gemm_f32(a: torch.Tensor, w: torch.Tensor, out: torch.Tensor, *, n: int | None = None, cta_group: int = 2) -> torch.Tensor:
    """out[M, n] (fp32, pitch out.stride(0)) = a[M,K] @ w[:n,K]^T.  `n` (multiple of 32) may exceed the rows the caller
    cares about when the memory behind w is valid (padded score columns of the VAE attention)."""
    _req(a, "a"); _req(w, "w"); _req(out, "out", torch.float32)
    M, K = a.shape
    N = w.shape[0] if n is None else n
    if out.shape[0] != M or out.shape[1] < N:
        raise ValueError("gemm_f32: out too small")
    _call("gemm", 2.0 * M * N * K, load().gf_gemm_bf16, ctx(), a.data_ptr(), _ld(a), w.data_ptr(), _ld(w), out.data_ptr(),
          _ld(out), M, N, K, None, GF_EPI_F32, None, None, 0, cta_group, _stream())
    return out


def conv3d_cl(x: torch.Tensor, w: torch.Tensor, bias: torch.Tensor | None, *, kernel, stride=(1, 1, 1), pad=(0, 0, 0),
              out_dims=None, out: torch.Tensor | None = None, residual: torch.Tensor | None = None,
              norm_out: torch.Tensor | None = None, gamma: torch.Tensor | None = None, silu: bool = True,
              cout: int | None = None, ncthw: bool = False, want_raw: bool = True):
    """x: [T, H, W, Cin] channels-last view (position pitch x.stride(2)); w: [Cout, taps*Cin]; see gf_conv3d_cl_bf16.
    Returns (Y or None, Y2 or None)."""
    _req(x, "x"); _req(w, "w")
    T, H, W, Cin = x.shape
    ldx = x.stride(2)
    if x.stride(1) != W * ldx or x.stride(0) != H * W * ldx:
        raise ValueError("conv3d_cl: x must be a dense [T, H, W] grid of rows")
    kt, kh, kw = kernel
    Cout = w.shape[0] if cout is None else cout
    if w.shape[1] != kt * kh * kw * Cin or not w.is_contiguous():
        raise ValueError(f"conv3d_cl: weight {tuple(w.shape)} does not match taps*Cin = {kt * kh * kw * Cin}")
    To, Ho, Wo = out_dims if out_dims is not None else (T, H, W)
    cs = (Cout + 7) // 8 * 8
    if ncthw:
        if out is None:
            out = torch.empty((Cout, To, Ho, Wo), dtype=torch.bfloat16, device=x.device)
        ldy = 0
    else:
        if out is None and want_raw:
            out = torch.empty((To, Ho, Wo, cs), dtype=torch.bfloat16, device=x.device)
        ldy = out.stride(2) if out is not None else 0
    if gamma is not None and norm_out is None:
        norm_out = torch.empty((To, Ho, Wo, cs), dtype=torch.bfloat16, device=x.device)
    for nme, t in (("bias", bias), ("residual", residual), ("gamma", gamma), ("out", out), ("norm_out", norm_out)):
        if t is not None:
            _req(t, nme)
    _call(f"conv3d:{Cin}>{Cout}:k{kt}{kh}{kw}s{stride[0]}{stride[1]}", 2.0 * To * Ho * Wo * Cout * kt * kh * kw * Cin,
          load().gf_conv3d_cl_bf16, ctx(), x.data_ptr(), ldx, T,
          H, W, Cin, w.data_ptr(), Cout, kt, kh, kw, stride[0], stride[1], stride[2], pad[0], pad[1], pad[2], _ptr(bias),
          _ptr(out), ldy, To, Ho, Wo, _ptr(residual), residual.stride(2) if residual is not None else 0,
          _ptr(norm_out), norm_out.stride(2) if norm_out is not None else 0, _ptr(gamma), 1 if silu else 0,
          1 if ncthw else 0, _stream())
    return out, norm_out


def vae_rmsnorm(x: torch.Tensor, gamma: torch.Tensor, *, silu: bool, out: torch.Tensor | None = None) -> torch.Tensor:
    """x: [..., C] rows with a uniform pitch (x.stride(-2)); RMS_norm over C (+ SiLU)."""
    _req(x, "x"); _req(gamma, "gamma")
    C = x.shape[-1]
    rows = x.numel() // C
    if out is None:
        out = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    _call("vae_rowwise", 4.0 * rows * C, load().gf_vae_rmsnorm_bf16, x.data_ptr(), x.stride(-2), out.data_ptr(),
          out.stride(-2), rows, C, gamma.data_ptr(), 1 if silu else 0, _stream())
    return out


def vae_upsample2x(first: torch.Tensor, rest: torch.Tensor | None, F: int, H: int, W: int, C: int,
                   out: torch.Tensor | None = None) -> torch.Tensor:
    _req(first, "first")
    if rest is not None:
        _req(rest, "rest")
    if out is None:
        out = torch.empty((F, 2 * H, 2 * W, C), dtype=torch.bfloat16, device=first.device)
    _call("vae_rowwise", 10.0 * F * H * W * C, load().gf_vae_upsample2x_bf16, first.data_ptr(), first.stride(-2),
          _ptr(rest), rest.stride(-2) if rest is not None else 0, out.data_ptr(), out.stride(-2), F, H, W, C, _stream())
    return out


def softmax_f32(S: torch.Tensor, P: torch.Tensor, L: int, Lp: int, scale: float) -> torch.Tensor:
    _req(S, "S", torch.float32); _req(P, "P")
    _call("vae_rowwise", 6.0 * S.shape[0] * L, load().gf_softmax_f32_bf16, S.data_ptr(), S.stride(0), P.data_ptr(),
          P.stride(0), S.shape[0], L, Lp, scale, _stream())
    return P


def vae_planes_to_cl(src: torch.Tensor, Cp: int, *, mean: torch.Tensor | None = None,
                     inv_std: torch.Tensor | None = None, out: torch.Tensor | None = None, wpad: int = 0) -> torch.Tensor:
    """src: (C, T, H, W) contiguous bf16 -> [T, H, W, Cp] channels-last (zero-padded channels).  wpad > 0: `out` is a
    zero-filled [T, H, W + 2*wpad, Cp] buffer and the clip lands in its columns [wpad, wpad + W)."""
    _req(src, "src")
    if not src.is_contiguous():
        raise ValueError("vae_planes_to_cl: src must be contiguous")
    C, T, H, W = src.shape
    if wpad and (out is None or out.shape != (T, H, W + 2 * wpad, Cp) or not out.is_contiguous()):
        raise ValueError("vae_planes_to_cl: wpad needs a contiguous [T, H, W + 2*wpad, Cp] destination")
    if out is None:
        out = torch.empty((T, H, W, Cp), dtype=torch.bfloat16, device=src.device)
    mode = 0 if mean is None else 1
    if mode:
        _req(mean, "mean", torch.float32); _req(inv_std, "inv_std", torch.float32)
    _call("vae_rowwise", 2.0 * T * H * W * (C + Cp), load().gf_vae_planes_to_cl_bf16, src.data_ptr(), T * H * W, C,
          out.data_ptr(), out.stride(2), Cp, _ptr(mean), _ptr(inv_std), mode, W if wpad else 0, wpad, _stream())
    return out


def vae_cl_to_planes(src: torch.Tensor, C: int, *, mean: torch.Tensor | None = None,
                     inv_std: torch.Tensor | None = None) -> torch.Tensor:
    """src: [T, H, W, >=C] channels-last -> (C, T, H, W) bf16."""
    _req(src, "src")
    T, H, W, _ = src.shape
    out = torch.empty((C, T, H, W), dtype=torch.bfloat16, device=src.device)
    mode = 0 if mean is None else 1
    if mode:
        _req(mean, "mean", torch.float32); _req(inv_std, "inv_std", torch.float32)
    _call("vae_rowwise", 4.0 * T * H * W * C, load().gf_vae_cl_to_planes_bf16, src.data_ptr(), src.stride(2), T * H * W, C,
          out.data_ptr(), _ptr(mean), _ptr(inv_std), mode, _stream())
    return out


def vae_head_gather(D: torch.Tensor, bias: torch.Tensor, C: int) -> torch.Tensor:
    """D: [T, H, W, >= 36] partial sums per spatial tap (channel tap*4 + co); bias fp32 [C] -> (C, T, H, W) bf16."""
    _req(D, "D"); _req(bias, "bias", torch.float32)
    T, H, W, _ = D.shape
    out = torch.empty((C, T, H, W), dtype=torch.bfloat16, device=D.device)
    _call("vae_rowwise", 2.0 * T * H * W * (D.shape[3] + C), load().gf_vae_head_gather_bf16, D.data_ptr(), D.stride(2),
          bias.data_ptr(), out.data_ptr(), C, T, H, W, _stream())
    return out


def vae_blend_(values: torch.Tensor, tile: torch.Tensor, mask: torch.Tensor, h0: int, w0: int) -> None:
    """values (C, T, H, W) += tile (C, T, th, tw) * mask (th, tw), reference rounding."""
    _req(values, "values"); _req(tile, "tile"); _req(mask, "mask")
    if not (values.is_contiguous() and tile.is_contiguous() and mask.is_contiguous()):
        raise ValueError("vae_blend_: contiguous tensors required")
    C, T, H, W = values.shape
    th, tw = tile.shape[2], tile.shape[3]
    _call("vae_rowwise", 6.0 * tile.numel(), load().gf_vae_blend_bf16, values.data_ptr(), C, T, H, W, tile.data_ptr(),
          th, tw, h0, w0, mask.data_ptr(), _stream())


def vae_blend_finish_(values: torch.Tensor, weight: torch.Tensor | None, clamp: bool) -> torch.Tensor:
    _req(values, "values")
    C, T, H, W = values.shape
    if weight is not None:
        _req(weight, "weight")
    _call("vae_rowwise", 4.0 * values.numel(), load().gf_vae_blend_finish_bf16, values.data_ptr(), C * T, H, W,
          _ptr(weight), 1 if clamp else 0, _stream())
    return values


# ------------------------------------------------------------------------------------------------ peer memory
GF_PEER_HANDLE_BYTES = 64


class _RawDeviceMemory:
    """__cuda_array_interface__ shim: lets torch view a device buffer owned by the library (no copy)."""

    def __init__(self, ptr: int, n_int16: int):
        self.__cuda_array_interface__ = {"shape": (n_int16,), "typestr": "<i2", "data": (ptr, False), "version": 3}


def peer_alloc(nbytes: int) -> int:
    """Zero-filled device buffer that other processes on the node can map (gf_peer_alloc). Returns the pointer."""
    out = ctypes.c_void_p()
    _check(load().gf_peer_alloc(ctypes.byref(out), nbytes), "gf_peer_alloc")
    return out.value


def peer_free(ptr: int) -> None:
    _check(load().gf_peer_free(ptr), "gf_peer_free")


def peer_export(ptr: int) -> bytes:
    buf = ctypes.create_string_buffer(GF_PEER_HANDLE_BYTES)
    _check(load().gf_peer_export(ptr, buf), "gf_peer_export")
    return buf.raw


def peer_import(handle: bytes) -> int:
    out = ctypes.c_void_p()
    _check(load().gf_peer_import(handle, ctypes.byref(out)), "gf_peer_import")
    return out.value


def peer_unimport(ptr: int) -> None:
    _check(load().gf_peer_unimport(ptr), "gf_peer_unimport")


def as_bf16_tensor(ptr: int, shape, device) -> torch.Tensor:
    """bf16 tensor view of library-owned device memory."""
    n = 1
    for s in shape:
        n *= s
    t = torch.as_tensor(_RawDeviceMemory(ptr, n), device=device)
    return t.view(torch.bfloat16).view(*shape)


def ptr_array(ptrs) -> ctypes.Array:
    return (ctypes.c_void_p * len(ptrs))(*ptrs)


class PeerStatus:
    """Host-mapped status word written by gf_peer_barrier when a peer does not arrive in time (gf_peer_status_alloc).
    `check()` reads it from the CPU without synchronising the stream."""

    def __init__(self):
        h, d = ctypes.c_void_p(), ctypes.c_void_p()
        _check(load().gf_peer_status_alloc(ctypes.byref(h), ctypes.byref(d)), "gf_peer_status_alloc")
        self.host, self.dev = h.value, d.value
        self._word = ctypes.cast(self.host, ctypes.POINTER(ctypes.c_uint))

    def value(self) -> int:
        return int(self._word[0]) if self.host else 0

    def check(self, what: str = "gf_peer_barrier") -> None:
        v = self.value()
        if v:
            raise RuntimeError(f"{what}: peer rank {v - 1} of the sequence-parallel group did not reach the barrier "
                               f"within the timeout (GF_PEER_TIMEOUT_MS); results since then are invalid")

    def free(self) -> None:
        if self.host:
            load().gf_peer_status_free(self.host)
            self.host = self.dev = None


def peer_barrier(flag_ptrs: ctypes.Array, n_peers: int, rank: int, timeout_ms: int = 0,
                 status: PeerStatus | None = None) -> None:
    _call("peer_barrier", 0.0, load().gf_peer_barrier, flag_ptrs, n_peers, rank, timeout_ms,
          status.dev if status is not None else None, _stream())


def qkv_rmsnorm_rope_scatter(qkv: torch.Tensor, weight_q: torch.Tensor, weight_k: torch.Tensor, *, eps: float,
                             cos_sin: torch.Tensor, head_dim: int, recv_ptrs: ctypes.Array, n_peers: int, rank: int,
                             ld_recv: int) -> None:
    """q/k RMSNorm + RoPE and v, stored straight into every head owner's receive buffer (Ulysses send, fused)."""
    _req(qkv, "qkv"); _req(weight_q, "weight_q"); _req(weight_k, "weight_k"); _req(cos_sin, "cos_sin", torch.float32)
    rows, d = qkv.shape[0], weight_q.numel()
    if cos_sin.shape != (rows, head_dim // 2, 2) or not cos_sin.is_contiguous():
        raise ValueError(f"cos_sin must be contiguous [rows, head_dim/2, 2], got {tuple(cos_sin.shape)}")
    _call("rmsnorm_rope", 12.0 * rows * d, load().gf_qkv_rmsnorm_rope_scatter_bf16, qkv.data_ptr(), _ld(qkv), rows, d,
          weight_q.data_ptr(), weight_k.data_ptr(), eps, cos_sin.data_ptr(), head_dim, recv_ptrs, n_peers, rank,
          ld_recv, _stream())


def attention_scatter(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, heads: int, *, out_ptrs: ctypes.Array,
                      n_peers: int, ldo: int, rows_per_peer: int, col_offset: int, scale: float | None = None,
                      kv_len: int | None = None) -> None:
    """attention() whose output rows are stored into their owners' [rows_per_peer, ldo] buffers (Ulysses return)."""
    _req(q, "q"); _req(k, "k"); _req(v, "v")
    head_dim = 128
    Lq, Lk = q.shape[0], k.shape[0]
    if kv_len is not None:
        if not 0 < kv_len <= Lk:
            raise ValueError("kv_len must be in (0, number of key rows]")
        Lk = kv_len
    if scale is None:
        scale = head_dim ** -0.5
    _call("attention_self", 4.0 * Lq * Lk * heads * head_dim, load().gf_attention_scatter_bf16, ctx(), q.data_ptr(),
          _ld(q),
          k.data_ptr(), _ld(k), v.data_ptr(), _ld(v), out_ptrs, n_peers, ldo, rows_per_peer, col_offset, Lq, Lk, heads,
          head_dim, scale, _stream())
