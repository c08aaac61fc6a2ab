"""Wan2.1 video VAE (z_dim 16) on the B200 library -- host-side mirror of the reference's `WanVideoVAE`
(diffsynth/models/wan_video_vae.py:1057-1250; call sites src/goal_force/wan_video_new.py:733,779,801-803,855,912).

Same public surface (`encode`, `decode`, `single_*`, `tiled_*`, `build_mask`), same tiling and blending arithmetic, but
the network runs as one pass over the whole clip in channels-last layout instead of the reference's chunk-by-chunk walk
with a feature cache (the two are the same function: oracle/wan_vae_oracle.py, checked against the reference by
oracle/gen_golden.py):

  * every CausalConv3d / Conv2d is `gf_conv3d_cl_bf16` (implicit GEMM, tcgen05; csrc/gf_conv.cu) with the residual add
    and the NEXT layer's RMS_norm + SiLU fused into its epilogue whenever the channel row fits one tile (<= 256);
  * 1x1x1 shortcuts and the attention block's projections are `gf_gemm_bf16`; the single-head attention (head dim =
    channels) is S = q k^T (fp32 epilogue), `gf_softmax_f32_bf16`, P v -- three GEMM-shaped launches per frame;
  * the rest (norms that cannot fuse, 2x upsampling + frame interleave, layout changes, tile blending) are the
    streaming kernels of csrc/gf_vae.cu.

There is no PyTorch fallback: every op goes through goal_force_b200.capi and raises when the library is missing.
"""
from __future__ import annotations

import torch

from . import capi

MEAN = [-0.7571, -0.7089, -0.9113, 0.1075, -0.1745, 0.9653, -0.1517, 1.5508, 0.4134, -0.0715, 0.5517, -0.3632,
        -0.1922, -0.9497, 0.2503, -0.2921]
STD = [2.8184, 1.4541, 2.3275, 2.6558, 1.2196, 1.7708, 2.6052, 2.0743, 3.2687, 2.1526, 2.8652, 1.5579, 1.6382,
       1.1253, 2.8251, 1.9160]

FUSE_MAX_CHANNELS = 256      # gf_conv3d_cl_bf16 fuses the next RMS_norm when the whole channel row is in one tile


def _round8(c: int) -> int:
    return (c + 7) // 8 * 8


# Weight re-layouts (pure functions of the reference's tensors; tests/test_vae_cpu.py checks each against F.conv3d by
# emulating the kernels' GEMM view in torch).
def conv_weight_2d(weight: torch.Tensor) -> torch.Tensor:
    """Conv3d / Conv2d weight (Cout, Cin, [kt,] kh, kw) -> the implicit GEMM's K-major operand [Cout, taps * CinP]:
    tap-major (dt, dh, dw), channels innermost, Cin zero-padded to a multiple of 8 (fp32; the caller casts)."""
    if weight.dim() == 4:                           # Conv2d -> kt = 1
        weight = weight.unsqueeze(2)
    cout, cin, kt, kh, kw = weight.shape
    w = weight.detach().to(torch.float32).permute(0, 2, 3, 4, 1)
    if _round8(cin) != cin:
        w = torch.nn.functional.pad(w, (0, _round8(cin) - cin))
    return w.reshape(cout, kt * kh * kw * _round8(cin))


def fold_input_conv_weight(weight: torch.Tensor) -> torch.Tensor:
    """encoder.conv1 (Cout, 3, 3, 3, 3) -> [Cout, 9 * 64]: tap (dt, dh), and inside a tap the 64-element window of 8
    positions x 8 channels that starts one position left of the output pixel; dw = 0..2 are its first three positions."""
    cout, cin = weight.shape[0], weight.shape[1]
    wf = torch.zeros((cout, 3, 3, 8, 8), dtype=torch.float32, device=weight.device)        # [co, dt, dh, dw, c]
    wf[:, :, :, :3, :cin] = weight.detach().to(torch.float32).permute(0, 2, 3, 4, 1)
    return wf.reshape(cout, 9 * 64)


def head_tap_weight(weight: torch.Tensor) -> torch.Tensor:
    """decoder.head.2 (3, C, 3, 3, 3) -> [36, 3 * C]: output channel tap*4 + co (tap = dh*3 + dw, the 4th channel of a
    tap is zero), K = (dt, ci) -- the (3,1,1) convolution whose nine shifted partial sums the gather kernel adds."""
    co, ci = weight.shape[0], weight.shape[1]
    wh = torch.zeros((9, 4, 3, ci), dtype=torch.float32, device=weight.device)             # [tap, co, dt, ci]
    wh[:, :co] = weight.detach().to(torch.float32).permute(3, 4, 0, 2, 1).reshape(9, co, 3, ci)
    return wh.reshape(36, 3 * ci)


class _Conv:
    """One convolution's device-side weights: w2d [Cout, taps * CinP] bf16 (tap-major, channels innermost)."""

    __slots__ = ("w", "bias", "kernel", "cin", "cout")

    def __init__(self, weight: torch.Tensor, bias: torch.Tensor, device):
        k = tuple(weight.shape[2:]) if weight.dim() == 5 else (1,) + tuple(weight.shape[2:])
        cout, cin = weight.shape[0], weight.shape[1]
        self.w = conv_weight_2d(weight.to(device)).to(torch.bfloat16).contiguous()
        b = torch.zeros(_round8(cout), dtype=torch.bfloat16, device=device)
        b[:cout] = bias.detach().to(device=device, dtype=torch.bfloat16)
        self.bias = b
        self.kernel = k
        self.cin, self.cout = _round8(cin), cout


class WanVideoVAEB200:
    """Drop-in for `WanVideoVAE` on one B200.  `state_dict` uses the reference's key names (VideoVAE_.state_dict())."""

    upsampling_factor = 8

    def __init__(self, state_dict: dict, dim: int = 96, z_dim: int = 16, dim_mult=(1, 2, 4, 4), num_res_blocks: int = 2,
                 temperal_downsample=(False, True, True), device="cuda"):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("WanVideoVAEB200 needs a CUDA device (goal_force_b200 has no CPU path)")
        capi.load()
        self.dim, self.z_dim, self.dim_mult, self.num_res_blocks = dim, z_dim, tuple(dim_mult), num_res_blocks
        self.temporal_down = tuple(temperal_downsample)
        self.temporal_up = tuple(temperal_downsample[::-1])
        self.mean = torch.tensor(MEAN[:z_dim], dtype=torch.float32, device=self.device)
        self.inv_std = (1.0 / torch.tensor(STD[:z_dim])).to(dtype=torch.float32, device=self.device)
        self.conv: dict[str, _Conv] = {}
        self.gamma: dict[str, torch.Tensor] = {}
        self.attn: dict[str, dict] = {}
        for k, v in state_dict.items():
            if k.endswith(".gamma"):
                self.gamma[k[:-6]] = v.detach().reshape(-1).to(device=self.device, dtype=torch.bfloat16).contiguous()
            elif k.endswith(".weight") and ".to_qkv" not in k and ".proj" not in k:
                name = k[:-7]
                self.conv[name] = _Conv(v, state_dict[name + ".bias"], self.device)
        for k in state_dict:
            if k.endswith(".to_qkv.weight"):
                pre = k[: -len(".to_qkv.weight")]
                c = state_dict[k].shape[1]
                wqkv = state_dict[k].detach().reshape(3 * c, c).to(device=self.device, dtype=torch.bfloat16)
                bqkv = state_dict[pre + ".to_qkv.bias"].detach().to(device=self.device, dtype=torch.bfloat16)
                self.attn[pre] = {
                    "c": c, "wqk": wqkv[: 2 * c].contiguous(), "bqk": bqkv[: 2 * c].contiguous(),
                    "wv": wqkv[2 * c:].contiguous(), "bv": bqkv[2 * c:].contiguous(),
                    "wo": state_dict[pre + ".proj.weight"].detach().reshape(c, c).to(device=self.device,
                                                                                    dtype=torch.bfloat16).contiguous(),
                    "bo": state_dict[pre + ".proj.bias"].detach().to(device=self.device, dtype=torch.bfloat16).contiguous(),
                }
        self._masks: dict = {}
        # encoder.conv1 reads 3 (padded to 8) channels: as a 27-tap convolution every k-block would carry 8 useful
        # channels of 64.  Folded form: the clip is stored with one zero position left and right of every row, and a
        # position's operand row is the 64-element window [w, w + 8) x 8 channels of that padded row (position pitch 8
        # < Cin 64: overlapping windows).  The three dw taps become columns 0..23 of a (dt, dh)-tap weight, the rest
        # of the window meets zero weights: 9 k-blocks of 64 instead of 27.
        self.conv1_folded = fold_input_conv_weight(state_dict["encoder.conv1.weight"].to(self.device)).to(
            torch.bfloat16).contiguous()
        # decoder.head.2 (3x3x3 to 3 channels): N = 3 would waste the tensor core, so its nine spatial taps become output
        # channels of a (3,1,1) convolution (channel tap*4 + co) and gf_vae_head_gather_bf16 sums them per pixel
        self.head_taps = head_tap_weight(state_dict["decoder.head.2.weight"].to(self.device)).to(
            torch.bfloat16).contiguous()
        self.head_zero_bias = torch.zeros(40, dtype=torch.bfloat16, device=self.device)
        self.head_bias = state_dict["decoder.head.2.bias"].detach().to(device=self.device, dtype=torch.bfloat16).float()

    @classmethod
    def from_reference(cls, vae, device="cuda"):
        """`vae`: the reference's WanVideoVAE (or its VideoVAE_ `.model`)."""
        model = getattr(vae, "model", vae)
        dim = model.encoder.conv1.weight.shape[0]
        return cls(model.state_dict(), dim=dim, z_dim=model.z_dim, device=device)

    # ------------------------------------------------------------------------------------------------ layers
    def _conv3d(self, name, x, *, pad, stride=(1, 1, 1), out_dims=None, out=None, residual=None, next_norm=None,
                want_raw=True, ncthw=False, cout=None):
        """next_norm: (gamma, silu) of the consumer's RMS_norm.  Returns (raw or None, normed or None)."""
        cv = self.conv[name]
        cout = cv.cout if cout is None else cout
        # The consumer's RMS_norm + SiLU is fused whenever the channel row fits one tile: measured on the final kernels
        # (tools/conv_bench.py --epi rawnorm) it adds 1.3 ms to the 192 -> 96 resample convolution at 81 x 480 x 832 and
        # 0.3 ms to 384 -> 192, against 2-5 ms for a separate pass over the same tensor.
        fuse = next_norm is not None and cout <= FUSE_MAX_CHANNELS and not ncthw
        w = cv.w if cout == cv.cout else cv.w[:cout]
        y, yn = capi.conv3d_cl(x, w, cv.bias, kernel=cv.kernel, stride=stride, pad=pad, out_dims=out_dims, out=out,
                               residual=residual, gamma=next_norm[0] if fuse else None,
                               silu=next_norm[1] if fuse else True, cout=cout, ncthw=ncthw,
                               want_raw=want_raw or not fuse)
        if next_norm is not None and not fuse and not ncthw:
            yn = capi.vae_rmsnorm(y, next_norm[0], silu=next_norm[1])
        return y, yn

    def _resblock(self, pre, x, xn, next_norm):
        """ResidualBlock (:267-301).  x: raw input, xn: silu(norm(x)) if a producer already made it."""
        if xn is None:
            xn = capi.vae_rmsnorm(x, self.gamma[pre + ".residual.0"], silu=True)
        _, y1n = self._conv3d(pre + ".residual.2", xn, pad=(2, 1, 1),
                              next_norm=(self.gamma[pre + ".residual.3"], True), want_raw=False)
        del xn
        if (pre + ".shortcut") in self.conv:
            sc = self.conv[pre + ".shortcut"]
            T, H, W, C = x.shape
            h = capi.gemm(x.view(T * H * W, C), sc.w, sc.bias[: sc.cout]).view(T, H, W, sc.cout)
        else:
            h = x
        return self._conv3d(pre + ".residual.6", y1n, pad=(2, 1, 1), residual=h, next_norm=next_norm)

    def _attention(self, pre, x):
        """AttentionBlock (:304-342): per-frame single-head attention over the H*W positions, head dim = C."""
        a = self.attn[pre]
        T, H, W, C = x.shape
        L = H * W
        Lp = (L + 31) // 32 * 32
        rows = T * L
        xn = torch.zeros((rows + 32, C), dtype=torch.bfloat16, device=x.device)
        capi.vae_rmsnorm(x.view(rows, C), self.gamma[pre + ".norm"], silu=False, out=xn[:rows])
        qk = torch.empty((rows + 32, 2 * C), dtype=torch.bfloat16, device=x.device)
        capi.gemm(xn[:rows], a["wqk"], a["bqk"], out=qk[:rows])
        S = torch.empty((L, Lp), dtype=torch.float32, device=x.device)
        P = torch.empty((L, Lp), dtype=torch.bfloat16, device=x.device)
        Vt = torch.empty((C, Lp), dtype=torch.bfloat16, device=x.device)
        O = torch.empty((rows, C), dtype=torch.bfloat16, device=x.device)
        scale = float(C) ** -0.5
        for f in range(T):
            r0 = f * L
            capi.gemm(a["wv"], xn[r0:r0 + Lp], out=Vt)                       # V^T (bias added after P v: rows of P sum to 1)
            capi.gemm_f32(qk[r0:r0 + L, :C], qk[r0:r0 + Lp, C:], S, n=Lp)
            capi.softmax_f32(S, P, L, Lp, scale)
            capi.gemm(P, Vt, a["bv"], out=O[r0:r0 + L])
        y = capi.gemm(O, a["wo"], a["bo"], epi=capi.GF_EPI_GATE_RES, residual=x.view(rows, C))
        return y.view(T, H, W, C)

    def _upsample(self, pre, x, temporal, next_norm):
        """Resample 'upsample2d' / 'upsample3d' (:82-160)."""
        T, H, W, C = x.shape
        if temporal and T > 1:
            rest, _ = self._conv3d(pre + ".time_conv", x[1:], pad=(2, 0, 0))
            F = 2 * T - 1
            up = capi.vae_upsample2x(x, rest.view(T - 1, H, W, 2 * C), F, H, W, C)
            del rest
        else:
            F = T
            up = capi.vae_upsample2x(x, None, F, H, W, C)
        return self._conv3d(pre + ".resample.1", up, pad=(0, 1, 1), next_norm=next_norm)

    def _downsample(self, pre, x, temporal, next_norm):
        """Resample 'downsample2d' / 'downsample3d' (:82-174): ZeroPad2d((0,1,0,1)) + stride-2 Conv2d, then the strided
        (3,1,1) time_conv on frames (0,1,2), (2,3,4), ... with frame 0 passed through."""
        T, H, W, C = x.shape
        Ho, Wo = H // 2, W // 2
        if not (temporal and T > 1):
            return self._conv3d(pre + ".resample.1", x, pad=(0, 0, 0), stride=(1, 2, 2), out_dims=(T, Ho, Wo),
                                next_norm=next_norm)
        y, _ = self._conv3d(pre + ".resample.1", x, pad=(0, 0, 0), stride=(1, 2, 2), out_dims=(T, Ho, Wo))
        To = (T - 3) // 2 + 1
        out = torch.empty((1 + To, Ho, Wo, C), dtype=torch.bfloat16, device=x.device)
        out[0].copy_(y[0])
        self._conv3d(pre + ".time_conv", y, pad=(0, 0, 0), stride=(2, 1, 1), out_dims=(To, Ho, Wo), out=out[1:])
        yn = capi.vae_rmsnorm(out, next_norm[0], silu=next_norm[1]) if next_norm is not None else None
        return out, yn

    # ------------------------------------------------------------------------------------------------ networks
    def _decode_clip(self, z: torch.Tensor) -> torch.Tensor:
        """z: (16, T, h, w) bf16 on the device -> (3, 4T-3, 8h, 8w) bf16, un-clamped (VideoVAE_.decode :1011-1034)."""
        p = "decoder."
        x = capi.vae_planes_to_cl(z, _round8(self.z_dim), mean=self.mean, inv_std=self.inv_std)
        x, _ = self._conv3d("conv2", x, pad=(0, 0, 0))
        x, xn = self._conv3d(p + "conv1", x, pad=(2, 1, 1), next_norm=(self.gamma[p + "middle.0.residual.0"], True))
        x, _ = self._resblock(p + "middle.0", x, xn, None)
        x = self._attention(p + "middle.1", x)
        n_stage = len(self.dim_mult)
        per_stage = self.num_res_blocks + 1
        names = []                                   # (kind, name, temporal)
        idx = 0
        for i in range(n_stage):
            for _ in range(per_stage):
                names.append(("res", f"{p}upsamples.{idx}", False))
                idx += 1
            if i != n_stage - 1:
                names.append(("up", f"{p}upsamples.{idx}", self.temporal_up[i]))
                idx += 1
        seq = [("res", p + "middle.2", False)] + names
        xn = None
        for j, (kind, name, temporal) in enumerate(seq):
            if j + 1 < len(seq):
                nk, nn_, _ = seq[j + 1]
                next_norm = (self.gamma[nn_ + ".residual.0"], True) if nk == "res" else None
            else:
                next_norm = (self.gamma[p + "head.0"], True)
            if kind == "res":
                x, xn = self._resblock(name, x, xn, next_norm)
            else:
                x, xn = self._upsample(name, x, temporal, next_norm)
        del x
        part, _ = capi.conv3d_cl(xn, self.head_taps, self.head_zero_bias, kernel=(3, 1, 1), pad=(2, 0, 0))
        return capi.vae_head_gather(part, self.head_bias, self.conv[p + "head.2"].cout)

    def _encode_clip(self, video: torch.Tensor) -> torch.Tensor:
        """video: (3, T, H, W) bf16 on the device, T = 1 + 4k -> mu (16, 1+k, H/8, W/8) normalised
        (VideoVAE_.encode :984-1009)."""
        p = "encoder."
        n_stage = len(self.dim_mult)
        seq = []
        idx = 0
        for i in range(n_stage):
            for _ in range(self.num_res_blocks):
                seq.append(("res", f"{p}downsamples.{idx}", False))
                idx += 1
            if i != n_stage - 1:
                seq.append(("down", f"{p}downsamples.{idx}", self.temporal_down[i]))
                idx += 1
        seq.append(("res", p + "middle.0", False))
        C, T, H, W = video.shape
        padded = torch.zeros((T * H * (W + 2) + 8) * 8, dtype=torch.bfloat16, device=video.device)    # + window slack
        capi.vae_planes_to_cl(video, 8, out=padded[: T * H * (W + 2) * 8].view(T, H, W + 2, 8), wpad=1)
        windows = torch.as_strided(padded, (T, H, W + 2, 64), (H * (W + 2) * 8, (W + 2) * 8, 8, 1))
        cv = self.conv[p + "conv1"]
        # 9 k-blocks per tile: the MMAs are too short to hide a fused epilogue here (16.6 ms fused against 7.2 + 2.7 ms)
        x, _ = capi.conv3d_cl(windows, self.conv1_folded, cv.bias, kernel=(3, 3, 1), pad=(2, 1, 0), out_dims=(T, H, W))
        xn = capi.vae_rmsnorm(x, self.gamma[seq[0][1] + ".residual.0"], silu=True)
        del padded, windows
        for j, (kind, name, temporal) in enumerate(seq):
            if j + 1 < len(seq):
                nk, nn_, _ = seq[j + 1]
                next_norm = (self.gamma[nn_ + ".residual.0"], True) if nk == "res" else None
            else:
                next_norm = None                         # the attention block applies its own (SiLU-free) norm
            if kind == "res":
                x, xn = self._resblock(name, x, xn, next_norm)
            else:
                x, xn = self._downsample(name, x, temporal, next_norm)
        x = self._attention(p + "middle.1", x)
        x, xn = self._resblock(p + "middle.2", x, None, (self.gamma[p + "head.0"], True))
        x, _ = self._conv3d(p + "head.2", xn, pad=(2, 1, 1))
        mu, _ = self._conv3d("conv1", x, pad=(0, 0, 0), cout=self.z_dim)      # only the mean half of (mu, log_var)
        return capi.vae_cl_to_planes(mu, self.z_dim, mean=self.mean, inv_std=self.inv_std)

    # ------------------------------------------------------------------------------------------------ reference surface
    def build_1d_mask(self, length, left_bound, right_bound, border_width):
        x = torch.ones((length,))
        if not left_bound:
            x[:border_width] = (torch.arange(border_width) + 1) / border_width
        if not right_bound:
            x[-border_width:] = torch.flip((torch.arange(border_width) + 1) / border_width, dims=(0,))
        return x

    def build_mask(self, data, is_bound, border_width):
        """WanVideoVAE.build_mask (:1094-1106): fp32 (1, 1, 1, H, W) blending ramp for a tile of `data`'s H x W."""
        H, W = data.shape[-2], data.shape[-1]
        h = self.build_1d_mask(H, is_bound[0], is_bound[1], border_width[0]).view(H, 1).expand(H, W)
        w = self.build_1d_mask(W, is_bound[2], is_bound[3], border_width[1]).view(1, W).expand(H, W)
        return torch.stack([h, w]).min(dim=0).values.view(1, 1, 1, H, W)

    def _mask_dev(self, tile, is_bound, border_width):
        """build_mask(...).to(bf16) on the device as an [H, W] plane (what the reference multiplies with), cached."""
        key = (tile.shape[-2], tile.shape[-1], tuple(is_bound), tuple(border_width))
        m = self._masks.get(key)
        if m is None:
            m = self.build_mask(tile, is_bound, border_width)[0, 0, 0].to(device=self.device, dtype=torch.bfloat16)
            m = self._masks[key] = m.contiguous()
        return m

    @staticmethod
    def _tasks(H, W, size, stride):
        tasks = []
        for h in range(0, H, stride[0]):
            if h - stride[0] >= 0 and h - stride[0] + size[0] >= H:
                continue
            for w in range(0, W, stride[1]):
                if w - stride[1] >= 0 and w - stride[1] + size[1] >= W:
                    continue
                tasks.append((h, h + size[0], w, w + size[1]))
        return tasks

    def _prep(self, t: torch.Tensor) -> torch.Tensor:
        return t.to(device=self.device, dtype=torch.bfloat16).contiguous()

    def _tiles(self, tasks, fn, shape_of, group):
        """Yield (task, tile) in the reference's task order.  With a process group the tiles are computed round-robin
        by its ranks and broadcast from their owners (NCCL over NVLink), so every rank blends the same tiles in the
        same order and ends up with the bit-identical result of the single-GPU loop."""
        if group is None:
            for task in tasks:
                yield task, fn(task)
            return
        import torch.distributed as dist
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        mine = {i: fn(task) for i, task in enumerate(tasks) if i % world == rank}
        for i, task in enumerate(tasks):
            tile = mine.pop(i, None)
            if tile is None:
                tile = torch.empty(shape_of(task), dtype=torch.bfloat16, device=self.device)
            dist.broadcast(tile, src=dist.get_global_rank(group, i % world), group=group)
            yield task, tile

    def tiled_decode(self, hidden_states, device=None, tile_size=(34, 34), tile_stride=(18, 16), group=None):
        """`group`: optional torch.distributed process group whose ranks all hold `hidden_states` (the replica group
        after denoising); the tiles are then decoded round-robin across its GPUs."""
        _, _, T, H, W = hidden_states.shape
        up = self.upsampling_factor
        z = self._prep(hidden_states[0])
        out_T = T * 4 - 3
        values = torch.zeros((3, out_T, H * up, W * up), dtype=torch.bfloat16, device=self.device)
        weight = torch.zeros((1, 1, H * up, W * up), dtype=torch.bfloat16, device=self.device)
        border = ((tile_size[0] - tile_stride[0]) * up, (tile_size[1] - tile_stride[1]) * up)
        ones = {}
        tasks = self._tasks(H, W, tile_size, tile_stride)

        def shape_of(task):
            h, h_, w, w_ = task
            return (3, out_T, (min(h_, H) - h) * up, (min(w_, W) - w) * up)

        def decode_tile(task):
            h, h_, w, w_ = task
            return self._decode_clip(z[:, :, h:h_, w:w_].contiguous())

        for (h, h_, w, w_), tile in self._tiles(tasks, decode_tile, shape_of, group):
            mask = self._mask_dev(tile, (h == 0, h_ >= H, w == 0, w_ >= W), border)
            capi.vae_blend_(values, tile, mask, h * up, w * up)
            one = ones.setdefault(tuple(mask.shape), torch.ones((1, 1) + tuple(mask.shape), dtype=torch.bfloat16,
                                                                device=self.device))
            capi.vae_blend_(weight, one, mask, h * up, w * up)
        capi.vae_blend_finish_(values, weight.view(H * up, W * up), clamp=True)
        return values.unsqueeze(0)

    def tiled_encode(self, video, device=None, tile_size=(272, 272), tile_stride=(144, 128), group=None):
        """tile sizes in pixels (the public `encode` multiplies the latent-unit arguments by 8, as the reference)."""
        _, _, T, H, W = video.shape
        up = self.upsampling_factor
        v = self._prep(video[0])
        out_T = (T + 3) // 4
        values = torch.zeros((self.z_dim, out_T, H // up, W // up), dtype=torch.bfloat16, device=self.device)
        weight = torch.zeros((1, 1, H // up, W // up), dtype=torch.bfloat16, device=self.device)
        border = ((tile_size[0] - tile_stride[0]) // up, (tile_size[1] - tile_stride[1]) // up)
        ones = {}
        tasks = self._tasks(H, W, tile_size, tile_stride)

        def shape_of(task):
            h, h_, w, w_ = task
            return (self.z_dim, out_T, (min(h_, H) - h) // up, (min(w_, W) - w) // up)

        def encode_tile(task):
            h, h_, w, w_ = task
            return self._encode_clip(v[:, :, h:h_, w:w_].contiguous())

        for (h, h_, w, w_), tile in self._tiles(tasks, encode_tile, shape_of, group):
            mask = self._mask_dev(tile, (h == 0, h_ >= H, w == 0, w_ >= W), border)
            capi.vae_blend_(values, tile, mask, h // up, w // up)
            one = ones.setdefault(tuple(mask.shape), torch.ones((1, 1) + tuple(mask.shape), dtype=torch.bfloat16,
                                                                device=self.device))
            capi.vae_blend_(weight, one, mask, h // up, w // up)
        capi.vae_blend_finish_(values, weight.view(H // up, W // up), clamp=False)
        return values.unsqueeze(0)

    def single_encode(self, video, device=None):
        return self._encode_clip(self._prep(video[0])).unsqueeze(0)

    def single_decode(self, hidden_state, device=None):
        video = self._decode_clip(self._prep(hidden_state[0]))
        capi.vae_blend_finish_(video, None, clamp=True)
        return video.unsqueeze(0)

    def encode(self, videos, device=None, tiled=False, tile_size=(34, 34), tile_stride=(18, 16), group=None):
        hidden_states = []
        for video in videos:
            video = video.unsqueeze(0)
            if tiled:
                ts = (tile_size[0] * self.upsampling_factor, tile_size[1] * self.upsampling_factor)
                st = (tile_stride[0] * self.upsampling_factor, tile_stride[1] * self.upsampling_factor)
                hidden_state = self.tiled_encode(video, device, ts, st, group=group)
            else:
                hidden_state = self.single_encode(video, device)
            hidden_states.append(hidden_state.squeeze(0))
        return torch.stack(hidden_states)

    def decode(self, hidden_states, device=None, tiled=False, tile_size=(34, 34), tile_stride=(18, 16), group=None):
        videos = []
        for hidden_state in hidden_states:
            hidden_state = hidden_state.unsqueeze(0)
            if tiled:
                video = self.tiled_decode(hidden_state, device, tile_size, tile_stride, group=group)
            else:
                video = self.single_decode(hidden_state, device)
            videos.append(video.squeeze(0))
        return torch.stack(videos)
