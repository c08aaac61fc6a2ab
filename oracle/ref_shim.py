"""Import the UNMODIFIED reference (brown-palm/goal-force) from /root/reference on CPU -- TEST INFRASTRUCTURE ONLY.

Used by oracle/gen_golden.py (to produce tests/golden/) and by CPU tests that are skipped when /root/reference is
absent (it does not exist on the GPU box). Nothing is copied: the reference modules are imported where they lie.

The reference cannot be imported as-is in this image: `modelscope`, `imageio`, `ftfy`,
`controlnet_aux` are missing, and flash-attn *is* installed, which would route attention to a CUDA-only kernel.
The shim registers empty stand-in modules for the missing imports and clears FLASH_ATTN_2_AVAILABLE so that
flash_attention() takes its F.scaled_dot_product_attention branch (diffsynth/models/wan_video_dit.py:55-60).
"""
from __future__ import annotations

import importlib
import importlib.machinery
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("GF_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "diffsynth"))


class _Anything(types.ModuleType):
    """Stand-in module: any attribute is a dummy class; enough for `from x import Y` at import time."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        obj = type(name, (), {"__init__": lambda self, *a, **k: None})
        setattr(self, name, obj)
        return obj


def _stub(name: str) -> None:
    if name in sys.modules:
        return
    try:
        importlib.import_module(name)
        return
    except Exception:  # noqa: BLE001 - missing or broken optional dependency
        pass
    parts = name.split(".")
    for i in range(1, len(parts) + 1):
        full = ".".join(parts[:i])
        if full not in sys.modules:
            m = _Anything(full)
            m.__path__ = []  # behave like a package
            m.__spec__ = importlib.machinery.ModuleSpec(full, None, is_package=True)
            sys.modules[full] = m
            if i > 1:
                setattr(sys.modules[".".join(parts[:i - 1])], parts[i - 1], m)


_LOADED = {}


def _import_with_stubs(name: str, max_stubs: int = 32):
    """import `name`; every time a third-party module is missing, register a stand-in for exactly that module and
    retry (only absent modules are ever stubbed; nothing that is installed gets shadowed)."""
    for _ in range(max_stubs):
        try:
            return importlib.import_module(name)
        except ModuleNotFoundError as e:
            missing = e.name
            if not missing or missing.startswith(("diffsynth", "src")):
                raise
            for k in [k for k in sys.modules if k.startswith(("diffsynth", "src.")) or k == "src"]:
                del sys.modules[k]   # drop half-imported reference modules before retrying
            _stub(missing)
    raise RuntimeError(f"could not import {name} from the reference")


def load():
    """Returns a namespace with the reference objects on the hot path."""
    if _LOADED:
        return _LOADED["ns"]
    if not available():
        raise RuntimeError(f"reference not found at {REFERENCE_ROOT}")
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    dit = _import_with_stubs("diffsynth.models.wan_video_dit")
    dit.FLASH_ATTN_2_AVAILABLE = False
    dit.FLASH_ATTN_3_AVAILABLE = False
    dit.SAGE_ATTN_AVAILABLE = False
    pipe_mod = _import_with_stubs("src.goal_force.wan_video_new")
    sched = _import_with_stubs("diffsynth.schedulers.flow_match")
    ns = types.SimpleNamespace(
        dit=dit, WanModel=dit.WanModel, DiTBlock=dit.DiTBlock,
        pipe_mod=pipe_mod, model_fn_wan_video=pipe_mod.model_fn_wan_video, ControlNet=pipe_mod.ControlNet,
        FlowMatchScheduler=sched.FlowMatchScheduler,
    )
    _LOADED["ns"] = ns
    return ns


def load_module(name: str):
    """Import one reference module by dotted name (e.g. diffsynth.models.wan_video_text_encoder)."""
    if not available():
        raise RuntimeError(f"reference not found at {REFERENCE_ROOT}")
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    return _import_with_stubs(name)


def load_dataset_module():
    if not available():
        raise RuntimeError(f"reference not found at {REFERENCE_ROOT}")
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    return _import_with_stubs("src.goal_force.unified_dataset")


def cfg_kwargs(cfg) -> dict:
    """DiTConfig -> WanModel kwargs (wan_video_dit.py:273-294)."""
    return dict(dim=cfg.dim, in_dim=cfg.in_dim, ffn_dim=cfg.ffn_dim, out_dim=cfg.out_dim, text_dim=cfg.text_dim,
                freq_dim=cfg.freq_dim, eps=cfg.eps, patch_size=tuple(cfg.patch_size), num_heads=cfg.num_heads,
                num_layers=cfg.num_layers, has_image_input=False, require_clip_embedding=False)
