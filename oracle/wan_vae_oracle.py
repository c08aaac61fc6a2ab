"""torch restatement of the Wan2.1 video VAE (z_dim 16) -- TEST INFRASTRUCTURE ONLY (see oracle/wan_dit_oracle.py).

Follows diffsynth/models/wan_video_vae.py: CausalConv3d :33-52, RMS_norm :55-70, Resample :82-174, ResidualBlock
:267-301, AttentionBlock :304-342, Encoder3d :517-617, Decoder3d :736-839, VideoVAE_.encode / .decode :984-1034,
WanVideoVAE.tiled_decode / tiled_encode / build_mask :1081-1204.

The reference walks the clip chunk by chunk (1 frame, then 4 frames at a time when encoding / one latent frame at a
time when decoding) and carries the last two input frames of every causal convolution in a feature cache.  This file
states the SAME function over the whole clip at once:
  * a cached CausalConv3d is exactly a causal convolution over the full sequence with two zero frames in front;
  * `upsample3d`: the first frame bypasses `time_conv` and is not doubled ('Rep'); frames 1.. go through a causal
    `time_conv` whose history starts at frame 1 (zeros before it), and each yields two frames;
  * `downsample3d`: frame 0 passes; output j >= 1 is the (3,1,1) kernel on input frames 2j-2, 2j-1, 2j.
oracle/gen_golden.py checks this formulation against the reference's own chunked encode / decode (CPU, random
weights) and commits golden vectors; the CUDA path implements the full-clip form.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

MEAN = [-0.7571, -0.7089, -0.9113, 0.1075, -0.1745, 0.9653, -0.1517, 1.5508, 0.4134, -0.0715, 0.5517, -0.3632,
        -0.1922, -0.9497, 0.2503, -0.2921]
STD = [2.8184, 1.4541, 2.3275, 2.6558, 1.2196, 1.7708, 2.6052, 2.0743, 3.2687, 2.1526, 2.8652, 1.5579, 1.6382,
       1.1253, 2.8251, 1.9160]


def causal_conv3d(x, w, b, stride=(1, 1, 1)):
    """CausalConv3d over a full clip: pad (kt-1)*... = 2*padding[0] zero frames in front, symmetric spatial padding."""
    kt, kh, kw = w.shape[2:]
    x = F.pad(x, ((kw - 1) // 2, (kw - 1) // 2, (kh - 1) // 2, (kh - 1) // 2, kt - 1, 0))
    return F.conv3d(x, w, b, stride=stride)


def rms_norm(x, gamma, dim=1):
    return F.normalize(x, dim=dim) * (x.shape[dim] ** 0.5) * gamma


def _conv(sd, name, x, **kw):
    return causal_conv3d(x, sd[name + ".weight"], sd[name + ".bias"], **kw)


def residual_block(sd, pre, x):
    h = _conv(sd, pre + ".shortcut", x) if (pre + ".shortcut.weight") in sd else x
    y = F.silu(rms_norm(x, sd[pre + ".residual.0.gamma"]))
    y = _conv(sd, pre + ".residual.2", y)
    y = F.silu(rms_norm(y, sd[pre + ".residual.3.gamma"]))
    y = _conv(sd, pre + ".residual.6", y)
    return y + h


def attention_block(sd, pre, x):
    b, c, t, h, w = x.shape
    y = x.permute(0, 2, 1, 3, 4).reshape(b * t, c, h, w)
    y = rms_norm(y, sd[pre + ".norm.gamma"])
    qkv = F.conv2d(y, sd[pre + ".to_qkv.weight"], sd[pre + ".to_qkv.bias"])
    q, k, v = qkv.reshape(b * t, 1, c * 3, -1).permute(0, 1, 3, 2).contiguous().chunk(3, dim=-1)
    y = F.scaled_dot_product_attention(q, k, v)
    y = y.squeeze(1).permute(0, 2, 1).reshape(b * t, c, h, w)
    y = F.conv2d(y, sd[pre + ".proj.weight"], sd[pre + ".proj.bias"])
    y = y.reshape(b, t, c, h, w).permute(0, 2, 1, 3, 4)
    return y + x


def _per_frame(x, fn):
    b, c, t, h, w = x.shape
    y = fn(x.permute(0, 2, 1, 3, 4).reshape(b * t, c, h, w))
    return y.reshape(b, t, *y.shape[1:]).permute(0, 2, 1, 3, 4)


def upsample(sd, pre, x, temporal):
    if temporal and x.shape[2] > 1:
        b, c, t, h, w = x.shape
        rest = _conv(sd, pre + ".time_conv", x[:, :, 1:])                  # causal history starts at frame 1
        rest = rest.reshape(b, 2, c, t - 1, h, w)
        rest = torch.stack((rest[:, 0], rest[:, 1]), 3).reshape(b, c, 2 * (t - 1), h, w)
        x = torch.cat([x[:, :, :1], rest], 2)

    def f(y):
        y = F.interpolate(y.float(), scale_factor=(2.0, 2.0), mode="nearest-exact").type_as(y)
        return F.conv2d(y, sd[pre + ".resample.1.weight"], sd[pre + ".resample.1.bias"], padding=1)
    return _per_frame(x, f)


def downsample(sd, pre, x, temporal):
    def f(y):
        return F.conv2d(F.pad(y, (0, 1, 0, 1)), sd[pre + ".resample.1.weight"], sd[pre + ".resample.1.bias"], stride=2)
    x = _per_frame(x, f)
    if temporal and x.shape[2] > 1:
        w, bias = sd[pre + ".time_conv.weight"], sd[pre + ".time_conv.bias"]
        rest = F.conv3d(x, w, bias, stride=(2, 1, 1))                      # windows (0,1,2), (2,3,4), ...
        x = torch.cat([x[:, :, :1], rest], 2)
    return x


def _stage_plan(dim, dim_mult, num_res_blocks, decoder):
    if decoder:
        dims = [dim * u for u in [dim_mult[-1]] + dim_mult[::-1]]
    else:
        dims = [dim * u for u in [1] + dim_mult]
    return dims


def decoder(sd, x, dim=96, dim_mult=(1, 2, 4, 4), num_res_blocks=2, temporal_upsample=(True, True, False)):
    """Decoder3d.forward over the whole latent clip (x is already conv2'd)."""
    dims = _stage_plan(dim, list(dim_mult), num_res_blocks, True)
    p = "decoder."
    x = _conv(sd, p + "conv1", x)
    x = residual_block(sd, p + "middle.0", x)
    x = attention_block(sd, p + "middle.1", x)
    x = residual_block(sd, p + "middle.2", x)
    idx = 0
    for i in range(len(dims) - 1):
        for _ in range(num_res_blocks + 1):
            x = residual_block(sd, f"{p}upsamples.{idx}", x)
            idx += 1
        if i != len(dim_mult) - 1:
            x = upsample(sd, f"{p}upsamples.{idx}", x, temporal_upsample[i])
            idx += 1
    x = F.silu(rms_norm(x, sd[p + "head.0.gamma"]))
    return _conv(sd, p + "head.2", x)


def encoder(sd, x, dim=96, dim_mult=(1, 2, 4, 4), num_res_blocks=2, temporal_downsample=(False, True, True)):
    dims = _stage_plan(dim, list(dim_mult), num_res_blocks, False)
    p = "encoder."
    x = _conv(sd, p + "conv1", x)
    idx = 0
    for i in range(len(dims) - 1):
        for _ in range(num_res_blocks):
            x = residual_block(sd, f"{p}downsamples.{idx}", x)
            idx += 1
        if i != len(dim_mult) - 1:
            x = downsample(sd, f"{p}downsamples.{idx}", x, temporal_downsample[i])
            idx += 1
    x = residual_block(sd, p + "middle.0", x)
    x = attention_block(sd, p + "middle.1", x)
    x = residual_block(sd, p + "middle.2", x)
    x = F.silu(rms_norm(x, sd[p + "head.0.gamma"]))
    return _conv(sd, p + "head.2", x)


def scale_pair(dtype, device, z_dim=16):
    mean = torch.tensor(MEAN[:z_dim], dtype=dtype, device=device)
    inv_std = (1.0 / torch.tensor(STD[:z_dim])).to(dtype=dtype, device=device)
    return mean, inv_std


def decode(sd, z, **kw):
    """VideoVAE_.decode (:1011-1034): z (B, 16, T, h, w) -> video (B, 3, 4T-3, 8h, 8w), un-clamped."""
    mean, inv_std = scale_pair(z.dtype, z.device, z.shape[1])
    z = z / inv_std.view(1, -1, 1, 1, 1) + mean.view(1, -1, 1, 1, 1)
    x = _conv(sd, "conv2", z)
    return decoder(sd, x, **kw)


def encode(sd, video, **kw):
    """VideoVAE_.encode (:984-1009): video (B, 3, 1+4k, H, W) -> mu (B, 16, 1+k, H/8, W/8), normalised."""
    out = encoder(sd, video, **kw)
    mu, _ = _conv(sd, "conv1", out).chunk(2, dim=1)
    mean, inv_std = scale_pair(mu.dtype, mu.device, mu.shape[1])
    return (mu - mean.view(1, -1, 1, 1, 1)) * inv_std.view(1, -1, 1, 1, 1)


# ------------------------------------------------------------------------------------------------ tiling (WanVideoVAE)
def build_1d_mask(length, left_bound, right_bound, border_width):
    x = torch.ones((length,))
    if not left_bound:
        x[:border_width] = (torch.arange(border_width) + 1) / border_width
    if not right_bound:
        x[-border_width:] = torch.flip((torch.arange(border_width) + 1) / border_width, dims=(0,))
    return x


def build_mask(H, W, is_bound, border_width):
    h = build_1d_mask(H, is_bound[0], is_bound[1], border_width[0]).view(H, 1).expand(H, W)
    w = build_1d_mask(W, is_bound[2], is_bound[3], border_width[1]).view(1, W).expand(H, W)
    return torch.stack([h, w]).min(dim=0).values.view(1, 1, 1, H, W)


def tile_tasks(H, W, size, stride):
    tasks = []
    for h in range(0, H, stride[0]):
        if h - stride[0] >= 0 and h - stride[0] + size[0] >= H:
            continue
        for w in range(0, W, stride[1]):
            if w - stride[1] >= 0 and w - stride[1] + size[1] >= W:
                continue
            tasks.append((h, h + size[0], w, w + size[1]))
    return tasks


def tiled_decode(sd, z, tile_size=(34, 34), tile_stride=(18, 16), **kw):
    """WanVideoVAE.tiled_decode (:1103-1153) with the blending done in z's dtype, as the reference does."""
    _, _, T, H, W = z.shape
    up = 8
    weight = torch.zeros((1, 1, 4 * T - 3, H * up, W * up), dtype=z.dtype, device=z.device)
    values = torch.zeros((1, 3, 4 * T - 3, H * up, W * up), dtype=z.dtype, device=z.device)
    for h, h_, w, w_ in tile_tasks(H, W, tile_size, tile_stride):
        out = decode(sd, z[:, :, :, h:h_, w:w_], **kw)
        mask = build_mask(out.shape[3], out.shape[4], (h == 0, h_ >= H, w == 0, w_ >= W),
                          ((tile_size[0] - tile_stride[0]) * up, (tile_size[1] - tile_stride[1]) * up)
                          ).to(dtype=z.dtype, device=z.device)
        th, tw = h * up, w * up
        values[:, :, :, th:th + out.shape[3], tw:tw + out.shape[4]] += out * mask
        weight[:, :, :, th:th + out.shape[3], tw:tw + out.shape[4]] += mask
    return (values / weight).clamp_(-1, 1)


def tiled_encode(sd, video, tile_size=(34 * 8, 34 * 8), tile_stride=(18 * 8, 16 * 8), **kw):
    """WanVideoVAE.tiled_encode (:1155-1204); tile sizes in pixels."""
    _, _, T, H, W = video.shape
    up = 8
    weight = torch.zeros((1, 1, (T + 3) // 4, H // up, W // up), dtype=video.dtype, device=video.device)
    values = torch.zeros((1, 16, (T + 3) // 4, H // up, W // up), dtype=video.dtype, device=video.device)
    for h, h_, w, w_ in tile_tasks(H, W, tile_size, tile_stride):
        out = encode(sd, video[:, :, :, h:h_, w:w_], **kw)
        mask = build_mask(out.shape[3], out.shape[4], (h == 0, h_ >= H, w == 0, w_ >= W),
                          ((tile_size[0] - tile_stride[0]) // up, (tile_size[1] - tile_stride[1]) // up)
                          ).to(dtype=video.dtype, device=video.device)
        th, tw = h // up, w // up
        values[:, :, :, th:th + out.shape[3], tw:tw + out.shape[4]] += out * mask
        weight[:, :, :, th:th + out.shape[3], tw:tw + out.shape[4]] += mask
    return values / weight


def random_state_dict(dim=96, z_dim=16, dim_mult=(1, 2, 4, 4), num_res_blocks=2, seed=0, dtype=torch.float32,
                      device="cpu"):
    """Seeded weights with the reference's key names (state_dict of VideoVAE_), scaled so activations stay O(1)."""
    g = torch.Generator("cpu").manual_seed(seed)
    sd = {}

    def conv(name, cout, cin, k):
        fan = cin * k[0] * k[1] * k[2] if len(k) == 3 else cin * k[0] * k[1]
        sd[name + ".weight"] = torch.randn(cout, cin, *k, generator=g) * fan ** -0.5
        sd[name + ".bias"] = torch.randn(cout, generator=g) * 0.02

    def gamma(name, c, images):
        shape = (c, 1, 1) if images else (c, 1, 1, 1)
        sd[name] = 1.0 + 0.1 * torch.randn(shape, generator=g)

    def resblock(pre, cin, cout):
        gamma(pre + ".residual.0.gamma", cin, False)
        conv(pre + ".residual.2", cout, cin, (3, 3, 3))
        gamma(pre + ".residual.3.gamma", cout, False)
        conv(pre + ".residual.6", cout, cout, (3, 3, 3))
        if cin != cout:
            conv(pre + ".shortcut", cout, cin, (1, 1, 1))

    def attn(pre, c):
        gamma(pre + ".norm.gamma", c, True)
        conv(pre + ".to_qkv", 3 * c, c, (1, 1))
        conv(pre + ".proj", c, c, (1, 1))

    mult = list(dim_mult)
    # encoder
    dims = [dim * u for u in [1] + mult]
    conv("encoder.conv1", dims[0], 3, (3, 3, 3))
    idx = 0
    tdown = (False, True, True)
    for i, (cin, cout) in enumerate(zip(dims[:-1], dims[1:])):
        for _ in range(num_res_blocks):
            resblock(f"encoder.downsamples.{idx}", cin, cout)
            cin = cout
            idx += 1
        if i != len(mult) - 1:
            conv(f"encoder.downsamples.{idx}.resample.1", cout, cout, (3, 3))
            if tdown[i]:
                conv(f"encoder.downsamples.{idx}.time_conv", cout, cout, (3, 1, 1))
            idx += 1
    resblock("encoder.middle.0", cout, cout)
    attn("encoder.middle.1", cout)
    resblock("encoder.middle.2", cout, cout)
    gamma("encoder.head.0.gamma", cout, False)
    conv("encoder.head.2", z_dim * 2, cout, (3, 3, 3))
    conv("conv1", z_dim * 2, z_dim * 2, (1, 1, 1))
    conv("conv2", z_dim, z_dim, (1, 1, 1))
    # decoder
    dims = [dim * u for u in [mult[-1]] + mult[::-1]]
    conv("decoder.conv1", dims[0], z_dim, (3, 3, 3))
    resblock("decoder.middle.0", dims[0], dims[0])
    attn("decoder.middle.1", dims[0])
    resblock("decoder.middle.2", dims[0], dims[0])
    idx = 0
    tup = (True, True, False)
    for i, (cin, cout) in enumerate(zip(dims[:-1], dims[1:])):
        if i in (1, 2, 3):
            cin = cin // 2
        for _ in range(num_res_blocks + 1):
            resblock(f"decoder.upsamples.{idx}", cin, cout)
            cin = cout
            idx += 1
        if i != len(mult) - 1:
            conv(f"decoder.upsamples.{idx}.resample.1", cout // 2, cout, (3, 3))
            if tup[i]:
                conv(f"decoder.upsamples.{idx}.time_conv", cout * 2, cout, (3, 1, 1))
            idx += 1
    gamma("decoder.head.0.gamma", cout, False)
    conv("decoder.head.2", 3, cout, (3, 3, 3))
    return {k: v.to(dtype=dtype, device=device) for k, v in sd.items()}
