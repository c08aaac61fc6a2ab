#!/usr/bin/env python
"""Phase timeline of gf_attn80_kernel on one SM (needs a library built with -DGF_A8_TRACE, see tools/build_variant.py).

Runs the config-2 self-attention shape once with tracing on, and prints, per kv block of CTA 0, when (SM clock,
relative to the first traced event) every softmax warp and both MMA issuers passed their phase boundaries.
"""
import ctypes
import json
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

from goal_force_b200 import capi  # noqa: E402

NJ = 32


def main():
    impl = int(os.environ.get("TRACE_IMPL", "80"))
    lib = capi.load()
    lib.gf_debug_attn_trace.argtypes = [ctypes.c_void_p]
    heads, d, L = 40, 5120, 32760
    qkv = torch.randn(L, 3 * d, device="cuda").bfloat16()
    o = torch.empty(L, d, device="cuda", dtype=torch.bfloat16)
    capi.attention_tuning(impl, 0)
    run = lambda: capi.attention(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], heads, out=o)  # noqa: E731
    run(); torch.cuda.synchronize()
    buf = torch.zeros(20 * NJ * 8, dtype=torch.int64, device="cuda")
    lib.gf_debug_attn_trace(buf.data_ptr())
    run(); torch.cuda.synchronize()
    lib.gf_debug_attn_trace(None)
    t = buf.cpu().view(20, NJ, 8)
    t0 = int(t[t > 0].min())
    rel = torch.where(t > 0, t - t0, torch.full_like(t, -1))
    out = {"impl": impl, "softmax": rel[:16].tolist(), "issuer": rel[17:19].tolist()}
    (ROOT / "gpurun_out").mkdir(exist_ok=True)
    (ROOT / "gpurun_out" / f"attn_trace_{impl}.json").write_text(json.dumps(out))
    names = ["enter", "S ready", "S in regs", "max done", "exp done", "xchg done", "p_free", "P stored"]
    # per-block period and phase durations, averaged over the traced window, per warp
    print(f"impl {impl}: softmax warps (tile = w // 8, half = (w // 4) % 2, quarter = w % 4)")
    for w in range(16):
        r = rel[w].double()
        per = (r[1:, 0] - r[:-1, 0]).mean().item()
        d_ = [(r[:, k + 1] - r[:, k]).mean().item() for k in range(7)]
        print(f" w{w:2d} period {per:7.1f} | " + " ".join(f"{names[k + 1]}:{d_[k]:6.1f}" for k in range(7))
              + f" | start offset vs w0 {(r[:, 0] - rel[0, :, 0].double()).mean().item():7.1f}")
    inames = ["K ready", "s_free", "QK issued", "V ready", "p_full", "PV issued"]
    for k, w in enumerate((17, 18)):
        r = rel[w].double()
        if (r < 0).any():
            print(f" issuer {k}: no trace"); continue
        per = (r[1:, 0] - r[:-1, 0]).mean().item()
        d_ = [(r[:, j + 1] - r[:, j]).mean().item() for j in range(5)]
        print(f" issuer{k} period {per:7.1f} | " + " ".join(f"{inames[j + 1]}:{d_[j]:6.1f}" for j in range(5))
              + f" | offset vs w0 {(r[:, 0] - rel[0, :, 0].double()).mean().item():7.1f}")
    # one block in detail
    jj = 10
    print(f"block {jj} absolute times (softmax warps 0,4,8,12; issuers):")
    for w in (0, 4, 8, 12):
        print(f" w{w:2d} " + " ".join(f"{int(x):7d}" for x in rel[w, jj]))
    for w in (17, 18):
        print(f" i{w - 17:2d} " + " ".join(f"{int(x):7d}" for x in rel[w, jj][:6]))


if __name__ == "__main__":
    main()
