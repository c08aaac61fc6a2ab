#!/usr/bin/env python
"""Minimal driver for `ncu --set full` on the self-attention kernel at config-2 shape (32760 tokens, 40 heads)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from goal_force_b200 import capi

L = int(sys.argv[1]) if len(sys.argv) > 1 else 32760
heads = int(sys.argv[2]) if len(sys.argv) > 2 else 40
d = heads * 128
qkv = torch.randn(L, 3 * d, device="cuda").bfloat16()
o = torch.empty(L, d, device="cuda", dtype=torch.bfloat16)
for _ in range(2):
    capi.attention(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], heads, out=o)
torch.cuda.synchronize()
print("done")
