"""Full-size Wan VAE timings on one B200: the pipeline's default tiled decode / encode
(81 x 480 x 832, tiles (30, 52) / (15, 26) in latent units) and the untiled forms, device-timed with CUDA events, next
to the eager-PyTorch (cuDNN, bf16) full-clip oracle on the same GPU.  Random-init weights, synthetic inputs.
Usage: python tools/bench_vae.py [--out gpurun_out/vae_bench.json] [--skip-eager] [--frames 81]"""
import argparse
import json
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def timed(fn, warmup=1, iters=2):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters, out


def conv_flops(vae, fn):
    """algorithmic conv/GEMM flops of one call, from the binding's launch records"""
    from goal_force_b200 import capi
    capi.STATS.reset(timing=True)
    fn()
    torch.cuda.synchronize()
    s = capi.STATS.summary()
    capi.STATS.reset(timing=False)
    return {k: {"launches": v["launches"], "ms": round(v["ms"], 3), "work": v["work"]} for k, v in s.items()}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/vae_bench.json")
    ap.add_argument("--skip-eager", action="store_true")
    ap.add_argument("--frames", type=int, default=81)
    ap.add_argument("--height", type=int, default=480)
    ap.add_argument("--width", type=int, default=832)
    ap.add_argument("--launch-list", action="store_true", help="one untimed single decode + encode (ncu launch-list pass)")
    a = ap.parse_args()
    from goal_force_b200.wan_vae import WanVideoVAEB200
    from oracle import wan_vae_oracle as V
    from oracle import wan_dit_oracle as O
    torch.manual_seed(0)
    sd = V.random_state_dict(dim=96, seed=0)
    vae = WanVideoVAEB200(sd, dim=96)
    T = (a.frames + 3) // 4
    h, w = a.height // 8, a.width // 8
    z = torch.randn(1, 16, T, h, w, device="cuda").to(torch.bfloat16)
    video = (torch.rand(1, 3, a.frames, a.height, a.width, device="cuda") * 2 - 1).to(torch.bfloat16)
    res = {"frames": a.frames, "height": a.height, "width": a.width, "gpu": torch.cuda.get_device_name(0)}
    out = Path(a.out)
    out.parent.mkdir(exist_ok=True)

    def dump():
        out.write_text(json.dumps(res, indent=1))

    def section(name, fn):
        try:
            t0 = time.time()
            ms, y = timed(fn)
            res[name] = {"ms": round(ms, 2), "shape": list(y.shape), "wall_s": round(time.time() - t0, 1)}
            print(name, res[name], flush=True)
            return y
        except Exception as e:  # noqa: BLE001 -- record and go on: one GPU visit has to yield every number it can
            res[name] = {"error": repr(e)[:300]}
            print(name, "FAILED", repr(e)[:300], flush=True)
            torch.cuda.empty_cache()
            return None
        finally:
            dump()

    if a.launch_list:
        vae.decode(z, "cuda", tiled=False)
        vae.encode(video, "cuda", tiled=False)
        torch.cuda.synchronize()
        return
    tiled_dec = section("ours_tiled_decode", lambda: vae.decode(z, "cuda", tiled=True, tile_size=(30, 52), tile_stride=(15, 26)))
    single_dec = section("ours_single_decode", lambda: vae.decode(z, "cuda", tiled=False))
    section("ours_tiled_encode", lambda: vae.encode(video, "cuda", tiled=True, tile_size=(30, 52), tile_stride=(15, 26)))
    single_enc = section("ours_single_encode", lambda: vae.encode(video, "cuda", tiled=False))
    try:
        res["ours_single_decode_kernels"] = conv_flops(vae, lambda: vae.decode(z, "cuda", tiled=False))
        res["ours_single_encode_kernels"] = conv_flops(vae, lambda: vae.encode(video, "cuda", tiled=False))
        for key in ("ours_single_decode_kernels", "ours_single_encode_kernels"):
            tot_ms = tot_work = 0.0
            for tag, k in res[key].items():
                if tag.startswith("conv3d"):
                    k["tflops"] = round(k["work"] / k["ms"] / 1e9, 1)
                    tot_ms += k["ms"]
                    tot_work += k["work"]
            res[key]["conv3d_total"] = {"ms": round(tot_ms, 2), "tflops": round(tot_work / tot_ms / 1e9, 1)}
    except Exception as e:  # noqa: BLE001
        res["kernels_error"] = repr(e)[:300]
    dump()
    del tiled_dec
    torch.cuda.empty_cache()
    if not a.skip_eager:
        sdb = {k: v.to(device="cuda", dtype=torch.bfloat16) for k, v in sd.items()}
        with torch.no_grad():
            zt = z[:, :, :, :30, :52].contiguous()
            section("eager_bf16_decode_one_tile_30x52", lambda: V.decode(sdb, zt))
            ours_tile = section("ours_decode_one_tile_30x52", lambda: vae._decode_clip(zt[0]))
            # parity at full tile size (81 x 240 x 416) against the fp32 oracle on this GPU: ours and the oracle's bf16 run
            try:
                torch.backends.cudnn.allow_tf32 = False
                torch.backends.cuda.matmul.allow_tf32 = False
                sdf = {k: v.to(device="cuda", dtype=torch.float32) for k, v in sd.items()}
                want = V.decode(sdf, zt.float())[0]
                res["tile_decode_rel_l2_vs_fp32_oracle"] = {
                    "ours": O.rel_l2(ours_tile.float(), want),
                    "oracle_bf16": O.rel_l2(V.decode(sdb, zt)[0].float(), want)}
                vt = video[:, :, :, :240, :416].contiguous()
                want_e = V.encode(sdf, vt.float())[0]
                res["tile_encode_rel_l2_vs_fp32_oracle"] = {
                    "ours": O.rel_l2(vae._encode_clip(vt[0]).float(), want_e),
                    "oracle_bf16": O.rel_l2(V.encode(sdb, vt)[0].float(), want_e)}
                print({k: v for k, v in res.items() if "fp32_oracle" in k}, flush=True)
                del sdf, want, want_e
            except Exception as e:  # noqa: BLE001
                res["tile_parity_error"] = repr(e)[:300]
            torch.cuda.empty_cache()
            dump()
            eager_full = section("eager_bf16_single_decode", lambda: V.decode(sdb, z))
            if eager_full is not None and single_dec is not None:
                # parity at full size: ours vs the eager bf16 run, relative to the eager run's own scale
                res["full_size_decode_rel_l2_ours_vs_eager_bf16"] = O.rel_l2(single_dec.float().clamp(-1, 1),
                                                                             eager_full.float().clamp(-1, 1))
            del eager_full, ours_tile
            torch.cuda.empty_cache()
            eager_enc = section("eager_bf16_single_encode", lambda: V.encode(sdb, video))
            if eager_enc is not None and single_enc is not None:
                res["full_size_encode_rel_l2_ours_vs_eager_bf16"] = O.rel_l2(single_enc.float(), eager_enc.float())
    dump()
    print(json.dumps(res))


if __name__ == "__main__":
    main()
