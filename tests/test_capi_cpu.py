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


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from goal_force_b200 import capi
    monkeypatch.setattr(capi, "_LIB", None)
    monkeypatch.setattr(capi, "lib_path", lambda: tmp_path / "nope.so")
    monkeypatch.delenv("GF_B200_AUTOBUILD", raising=False)
    with pytest.raises(RuntimeError, match="no CPU/PyTorch fallback"):
        capi.load()
