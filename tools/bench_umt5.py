#!/usr/bin/env python
"""umT5-XXL prompt encoder at full size on one GPU: 24 layers, dim 4096, 64 heads, ffn 10240, vocab
256,384 (11.3 GB of bf16 weights, random init), the positive and the negative prompt of one video as one batch of two
512-token sequences.  Reports time per call (CUDA events) and parity against the oracle (reference restatement) run on
the same GPU with the same weights in bf16 and fp32.

    python tools/bench_umt5.py > profiles/rNN_umt5.json
"""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402


def main():
    torch.backends.cuda.matmul.allow_tf32 = False
    from goal_force_b200.umt5 import UMT5Config, UMT5EncoderB200
    from oracle import umt5_oracle as U
    from oracle import wan_dit_oracle as O
    cfg = UMT5Config()
    dev = "cuda"
    g = torch.Generator(dev).manual_seed(0)
    rn = lambda *s, std=1.0: (torch.randn(*s, generator=g, device=dev, dtype=torch.bfloat16) * std)  # noqa: E731
    d, da, df = cfg.dim, cfg.dim_attn, cfg.dim_ffn
    sd = {"token_embedding.weight": rn(cfg.vocab, d), "norm.weight": 1.0 + rn(d, std=0.1)}
    for i in range(cfg.num_layers):
        p = f"blocks.{i}."
        sd[p + "norm1.weight"] = 1.0 + rn(d, std=0.1)
        sd[p + "norm2.weight"] = 1.0 + rn(d, std=0.1)
        sd[p + "attn.q.weight"] = rn(da, d, std=d ** -0.5 * 0.35)
        sd[p + "attn.k.weight"] = rn(da, d, std=d ** -0.5)
        sd[p + "attn.v.weight"] = rn(da, d, std=d ** -0.5)
        sd[p + "attn.o.weight"] = rn(d, da, std=da ** -0.5)
        sd[p + "ffn.gate.0.weight"] = rn(df, d, std=d ** -0.5)
        sd[p + "ffn.fc1.weight"] = rn(df, d, std=d ** -0.5)
        sd[p + "ffn.fc2.weight"] = rn(d, df, std=df ** -0.5)
        sd[p + "pos_embedding.embedding.weight"] = rn(cfg.num_buckets, cfg.num_heads, std=0.5)
    ids, mask = U.synthetic_prompt(cfg.vocab, 2, 512, (140, 96), seed=1)
    ids, mask = ids.to(dev), mask.to(dev)
    enc = UMT5EncoderB200(cfg, sd)
    out = enc(ids, mask)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        enc(ids, mask)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    kw = dict(num_heads=cfg.num_heads, num_layers=cfg.num_layers, num_buckets=cfg.num_buckets)
    with torch.no_grad():
        refbf = U.encoder(sd, ids, mask, **kw)
        e0.record()
        U.encoder(sd, ids, mask, **kw)
        e1.record()
        torch.cuda.synchronize()
        ms_ref = e0.elapsed_time(e1)
        del enc
        sd32 = {k: v.float() for k, v in sd.items()}
        ref32 = U.encoder(sd32, ids, mask, **kw)
    flops = 2 * 512 * cfg.num_layers * 2.0 * (4 * d * da + 3 * d * df) + 2 * cfg.num_layers * 4.0 * 512 * 512 * da
    print(json.dumps({"what": "umT5-XXL encoder, 24 layers, two 512-token prompts in one batch, bf16, random init",
                      "ms_per_call": ms, "tflops": flops / ms / 1e9, "eager_oracle_ms_per_call": ms_ref,
                      "rel_l2_ours_vs_fp32": O.rel_l2(out, ref32), "rel_l2_oracle_bf16_vs_fp32": O.rel_l2(refbf, ref32),
                      "rel_l2_ours_vs_oracle_bf16": O.rel_l2(out, refbf),
                      "pass": O.rel_l2(out, ref32) <= max(1e-2, O.rel_l2(refbf, ref32)),
                      "gpu": torch.cuda.get_device_name(0)}))


if __name__ == "__main__":
    main()
