"""C-ABI boundary checks that need no GPU: the library loads, exports every symbol include/goalforce_b200.h declares,
and the Python wrappers refuse CPU tensors (there is no CPU fallback)."""
import ctypes
import re
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent


def _declared_symbols():
    text = (ROOT / "include" / "goalforce_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\bint\s+(gf_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_expected_entry_points():
    syms = _declared_symbols()
    for s in ("gf_gemm_bf16", "gf_attention_bf16", "gf_layernorm_bf16", "gf_rmsnorm_rope_bf16",
              "gf_patch_gather_bf16", "gf_unpatchify_bf16", "gf_cfg_euler_bf16"):
        assert s in syms


def test_library_exports_every_declared_symbol(lib):
    raw = ctypes.CDLL(str(ROOT / "goal_force_b200" / "_lib" / "libgoalforce_b200.so"))
    for s in _declared_symbols():
        assert hasattr(raw, s), f"{s} declared in the header but not exported"


def test_python_binding_covers_header(lib):
    from goal_force_b200 import capi
    assert sorted(capi.SIGNATURES) == _declared_symbols()
    assert lib.gf_abi_version() == capi.ABI_VERSION
    text = (ROOT / "include" / "goalforce_b200.h").read_text()
    assert int(re.search(r"#define GF_ABI_VERSION (\d+)", text).group(1)) == capi.ABI_VERSION


def test_binding_arity_matches_header(lib):
    """capi.SIGNATURES and include/goalforce_b200.h stay in lockstep: same number of parameters per entry point."""
    from goal_force_b200 import capi
    text = re.sub(r"/\*.*?\*/", "", (ROOT / "include" / "goalforce_b200.h").read_text(), flags=re.S)
    for name, params in re.findall(r"\bint\s+(gf_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", text, flags=re.S):
        params = params.strip()
        n = 0 if params in ("", "void") else params.count(",") + 1
        assert len(capi.SIGNATURES[name]) == n, name


def test_context_is_per_object_not_process_wide(lib):
    """gf_ctx carries tuning and the descriptor cache; two contexts do not see each other's settings, and a NULL
    context is a valid stateless call (argument validation still comes first)."""
    a, b = ctypes.c_void_p(), ctypes.c_void_p()
    assert lib.gf_ctx_create(ctypes.byref(a)) == 0 and lib.gf_ctx_create(ctypes.byref(b)) == 0
    assert a.value and b.value and a.value != b.value
    assert lib.gf_ctx_set_attention(a, 128, 4) == 0
    assert lib.gf_ctx_set_attention(a, 96, 0) == -1 and lib.gf_ctx_set_attention(a, 80, 3) == -1
    assert lib.gf_ctx_set_gemm_raster(b, 16) == 0 and lib.gf_ctx_set_gemm_raster(b, -2) == -1
    assert lib.gf_ctx_set_gemm_tile(b, 224) == 0 and lib.gf_ctx_set_gemm_tile(b, 0) == 0 and lib.gf_ctx_set_gemm_tile(b, 192) == -1
    assert lib.gf_ctx_set_conv(b, 2) == 0 and lib.gf_ctx_set_conv(b, 3) == -1
    e, h, m = ctypes.c_longlong(7), ctypes.c_longlong(7), ctypes.c_longlong(7)
    assert lib.gf_ctx_stats(a, ctypes.byref(e), ctypes.byref(h), ctypes.byref(m)) == 0
    assert (e.value, h.value, m.value) == (0, 0, 0)
    assert lib.gf_ctx_set_attention(None, 80, 0) == -1
    assert lib.gf_ctx_destroy(a) == 0 and lib.gf_ctx_destroy(b) == 0 and lib.gf_ctx_destroy(None) == -1


def test_no_cpu_fallback():
    from goal_force_b200 import capi
    a = torch.zeros(4, 8, dtype=torch.bfloat16)
    with pytest.raises(ValueError, match="CUDA"):
        capi.gemm(a, a)
    with pytest.raises(ValueError, match="CUDA"):
        capi.layernorm(a, eps=1e-6, shift=a[0], scale=a[0])


def test_bad_arguments_are_rejected_without_a_gpu(lib):
    # argument validation happens before any CUDA call, so it can be exercised on a CPU box
    GF_ERR_BAD_ARG = -1
    assert lib.gf_gemm_bf16(None, None, 8, None, 8, None, 8, 4, 32, 8, None, 0, None, None, 0, 1, None) == GF_ERR_BAD_ARG
    assert lib.gf_attention_bf16(None, None, 8, None, 8, None, 8, None, 8, 1, 1, 1, 128, 1.0, None) == GF_ERR_BAD_ARG
    assert lib.gf_add_bf16(None, None, None, 8, None) == GF_ERR_BAD_ARG
    assert lib.gf_peer_barrier(None, 2, 0, 0, None, None) == GF_ERR_BAD_ARG
    assert lib.gf_cfg_euler_bf16(None, None, None, None, 1.0, 0.0, 4, None) == GF_ERR_BAD_ARG


def test_vae_entry_points_validate_arguments_without_a_gpu(lib):
    """The VAE entry points reject bad arguments before any CUDA call; with valid-looking arguments
    and no driver the convolution reports GF_ERR_NO_DRIVER (tensor-map encoding) instead of crashing."""
    BAD, UNSUP = -1, -4
    buf = ctypes.create_string_buffer(4096 + 64)
    ptr = (ctypes.addressof(buf) + 63) & ~63

    def conv(**kw):
        a = dict(X=ptr, ldx=16, T=2, H=4, W=4, Cin=16, Wt=ptr, Cout=16, kt=3, kh=3, kw=3, st=1, sh=1, sw=1, pt=2, ph=1,
                 pw=1, bias=None, Y=ptr, ldy=16, To=2, Ho=4, Wo=4, R=None, ldr=0, Y2=None, ldy2=0, gamma=None, silu=1,
                 ncthw=0)
        a.update(kw)
        return lib.gf_conv3d_cl_bf16(None, a["X"], a["ldx"], a["T"], a["H"], a["W"], a["Cin"], a["Wt"], a["Cout"],
                                     a["kt"], a["kh"], a["kw"], a["st"], a["sh"], a["sw"], a["pt"], a["ph"], a["pw"],
                                     a["bias"], a["Y"], a["ldy"], a["To"], a["Ho"], a["Wo"], a["R"], a["ldr"], a["Y2"],
                                     a["ldy2"], a["gamma"], a["silu"], a["ncthw"], None)

    assert conv(X=None) == BAD and conv(Y=None) == BAD            # no output at all (neither Y nor Y2)
    assert conv(Cin=12) == BAD and conv(ldx=12) == BAD           # channels / pitch must be multiples of 8
    assert conv(X=ptr + 2) == BAD                                # 16-byte alignment
    assert conv(ldy=8) == BAD                                    # output pitch below the channel count
    assert conv(kt=4) == UNSUP                                   # 4 x 3 x 3 = 36 taps > 27
    assert conv(sh=3, sw=3) == UNSUP and conv(sh=2, sw=1) == UNSUP
    assert conv(Y2=ptr) == BAD                                   # fused norm without gamma
    assert conv(Cout=384, Y2=ptr, ldy2=384, gamma=ptr, ldy=384) == UNSUP     # channel row spans two n-tiles
    assert conv(ncthw=1, R=ptr, ldr=16) == UNSUP
    assert conv() in (-2, -3) or conv() > 0                      # no driver / no device on this box: a clean error code
    assert lib.gf_vae_rmsnorm_bf16(ptr, 16, ptr, 16, 4, 12, ptr, 1, None) == BAD          # C % 8
    assert lib.gf_vae_rmsnorm_bf16(ptr, 16, ptr, 16, 4, 2048, ptr, 1, None) == UNSUP      # C > 1024
    assert lib.gf_vae_upsample2x_bf16(None, 8, None, 0, ptr, 8, 1, 2, 2, 8, None) == BAD
    assert lib.gf_softmax_f32_bf16(None, 8, ptr, 8, 1, 8, 8, 1.0, None) == BAD
    assert lib.gf_softmax_f32_bf16(ctypes.cast(ptr, ctypes.POINTER(ctypes.c_float)), 8, ptr, 8, 1, 8, 4, 1.0, None) == BAD
    assert lib.gf_vae_planes_to_cl_bf16(ptr, 8, 3, ptr, 8, 4, None, None, 0, 0, 0, None) == BAD       # Cp % 8
    assert lib.gf_vae_planes_to_cl_bf16(ptr, 8, 3, ptr, 8, 8, None, None, 1, 0, 0, None) == BAD       # mode 1 needs mean / std
    assert lib.gf_vae_planes_to_cl_bf16(ptr, 9, 3, ptr, 8, 8, None, None, 0, 4, 1, None) == BAD       # N % row_w
    assert lib.gf_vae_cl_to_planes_bf16(ptr, 4, 8, 8, ptr, None, None, 0, None) == BAD                # ld < C
    assert lib.gf_vae_blend_bf16(ptr, 3, 1, 4, 4, ptr, 4, 4, 1, 0, ptr, None) == BAD                  # tile outside the canvas
    assert lib.gf_vae_blend_finish_bf16(None, 1, 4, 4, None, 1, None) == BAD
    assert lib.gf_ctx_set_conv(None, 0) == BAD


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from goal_force_b200 import capi
    monkeypatch.setattr(capi, "_LIB", None)
    monkeypatch.setattr(capi, "lib_path", lambda: tmp_path / "nope.so")
    monkeypatch.delenv("GF_B200_AUTOBUILD", raising=False)
    with pytest.raises(RuntimeError, match="no CPU/PyTorch fallback"):
        capi.load()
