"""Wan video VAE, CPU side: the full-clip oracle restatement against the vectors the reference's own
chunk-by-chunk VideoVAE_ / WanVideoVAE produced (tests/golden/vae.pt, oracle/gen_golden.py::gen_vae)."""
import pytest
import torch

from oracle import wan_dit_oracle as O
from oracle import wan_vae_oracle as V


def _case(golden_dir):
    g = torch.load(golden_dir / "vae.pt", weights_only=False)
    sd = V.random_state_dict(dim=g["dim"], seed=g["weight_seed"])
    return g, sd


def test_oracle_matches_reference_vectors(golden_dir):
    g, sd = _case(golden_dir)
    dim = g["dim"]
    with torch.no_grad():
        # decode goldens are stored as fp16 (|x| <= ~3): 1e-3 covers the storage rounding
        assert O.rel_l2(V.decode(sd, g["z"], dim=dim), g["decode"].float()) < 1e-3
        assert O.rel_l2(V.encode(sd, g["video"], dim=dim), g["encode"]) < 1e-5
        assert O.rel_l2(V.tiled_decode(sd, g["z_big"], (4, 5), (3, 3), dim=dim), g["tiled_decode"].float()) < 1e-3
        video = V.tiled_decode(sd, g["z_big"], (4, 5), (3, 3), dim=dim)
        assert O.rel_l2(V.tiled_encode(sd, video, (32, 40), (24, 24), dim=dim), g["tiled_encode"]) < 1e-4


def test_oracle_matches_reference_vectors_at_the_shipped_width(golden_dir):
    """dim 96 (96 / 192 / 384 channels, the Wan2.1 VAE's width): decode and encode of a tiny clip produced by the
    reference's own VideoVAE_ (oracle/gen_golden.py::gen_vae)."""
    g = torch.load(golden_dir / "vae.pt", weights_only=False)
    sd = V.random_state_dict(dim=96, seed=g["weight_seed96"])
    with torch.no_grad():
        assert O.rel_l2(V.decode(sd, g["z96"]), g["decode96"].float()) < 1e-3           # stored as fp16
        assert O.rel_l2(V.encode(sd, g["decode96"].float().clamp(-1, 1)), g["encode96"]) < 2e-3


def test_shapes_and_causality(golden_dir):
    g, sd = _case(golden_dir)
    dim = g["dim"]
    z = g["z"]
    with torch.no_grad():
        full = V.decode(sd, z, dim=dim)
        assert full.shape == (1, 3, 4 * z.shape[2] - 3, 8 * z.shape[3], 8 * z.shape[4])
        # causal in time: the first latent frame alone decodes to the first video frame of the full clip
        first = V.decode(sd, z[:, :, :1], dim=dim)
        assert O.rel_l2(first, full[:, :, :1]) < 1e-5
        two = V.decode(sd, z[:, :, :2], dim=dim)
        assert O.rel_l2(two, full[:, :, :5]) < 1e-5
        enc = V.encode(sd, g["video"], dim=dim)
        assert enc.shape == (1, 16, (g["video"].shape[2] + 3) // 4, g["video"].shape[3] // 8, g["video"].shape[4] // 8)
        enc5 = V.encode(sd, g["video"][:, :, :5], dim=dim)
        assert O.rel_l2(enc5, enc[:, :, :2]) < 1e-5


def test_tile_plan_matches_reference_loop():
    # WanVideoVAE.tiled_decode (:1108-1116): the pipeline's (30, 52) / (15, 26) tiles on a 60 x 104 latent
    tasks = V.tile_tasks(60, 104, (30, 52), (15, 26))
    assert len(tasks) == 9 and tasks[0] == (0, 30, 0, 52) and tasks[-1] == (30, 60, 52, 104)
    assert V.tile_tasks(34, 34, (34, 34), (18, 16)) == [(0, 34, 0, 34)]
    m = V.build_mask(16, 24, (True, False, False, True), (8, 8))
    assert m.shape == (1, 1, 1, 16, 24) and float(m[0, 0, 0, 0, 23]) == 1.0 and float(m[0, 0, 0, 15, 23]) == 0.125
    assert float(m[0, 0, 0, 0, 0]) == 0.125


@pytest.mark.reference
def test_state_dict_keys_match_live_reference():
    from oracle import ref_shim
    vae = ref_shim.load_module("diffsynth.models.wan_video_vae")
    m = vae.VideoVAE_(dim=32, z_dim=16)
    sd = V.random_state_dict(dim=32)
    ref = m.state_dict()
    assert sorted(ref) == sorted(sd)
    for k in ref:
        assert ref[k].shape == sd[k].shape, k


def test_standin_module_exposes_the_reference_surface(golden_dir):
    """The nn.Module stand-in used by the GPU drop-in test has the attribute surface `WanVideoVAEB200.from_reference`
    reads (model.z_dim, model.encoder.conv1.weight, model.state_dict() with the reference's keys)."""
    from oracle.ref_standins import WanVideoVAEStandIn
    g, sd = _case(golden_dir)
    m = WanVideoVAEStandIn(sd)
    assert sorted(m.model.state_dict()) == sorted(sd)
    assert m.model.encoder.conv1.weight.shape[0] == g["dim"] and m.model.z_dim == 16 and m.upsampling_factor == 8
    for k, v in m.model.state_dict().items():
        assert torch.equal(v, sd[k])


@pytest.mark.reference
def test_standin_matches_live_reference_attributes():
    from oracle import ref_shim
    from oracle.ref_standins import WanVideoVAEStandIn
    vae = ref_shim.load_module("diffsynth.models.wan_video_vae")
    ref = vae.WanVideoVAE(z_dim=16)
    ref.model = vae.VideoVAE_(dim=32, z_dim=16)
    m = WanVideoVAEStandIn(V.random_state_dict(dim=32))
    assert sorted(m.model.state_dict()) == sorted(ref.model.state_dict())
    assert m.model.z_dim == ref.model.z_dim and m.upsampling_factor == ref.upsampling_factor
    assert m.model.encoder.conv1.weight.shape == ref.model.encoder.conv1.weight.shape


def _im2col(x, kernel, pad):
    """(Cin, T, H, W) -> [T*H*W, taps * Cin] rows in the kernels' K order (tap-major (dt, dh, dw), channels innermost),
    leading zero padding `pad` and trailing padding so that the output keeps the input's extent."""
    import torch.nn.functional as F
    kt, kh, kw = kernel
    C, T, H, W = x.shape
    xp = F.pad(x, (pad[2], kw - 1 - pad[2], pad[1], kh - 1 - pad[1], pad[0], kt - 1 - pad[0]))
    cols = []
    for dt in range(kt):
        for dh in range(kh):
            for dw in range(kw):
                cols.append(xp[:, dt:dt + T, dh:dh + H, dw:dw + W].permute(1, 2, 3, 0).reshape(T * H * W, C))
    return torch.cat(cols, dim=1)


def test_weight_relayouts_against_conv3d():
    """Host-side weight layouts of goal_force_b200.wan_vae, each checked by emulating the kernel's GEMM in torch:
    the generic [Cout, taps*Cin] operand, the folded 64-element window of encoder.conv1, the tap-channel head."""
    import torch.nn.functional as F
    from goal_force_b200.wan_vae import conv_weight_2d, fold_input_conv_weight, head_tap_weight
    g = torch.Generator().manual_seed(4)
    # generic: causal 3x3x3, Cin padded 3 -> 8
    x = torch.randn(3, 4, 6, 7, generator=g)
    w = torch.randn(10, 3, 3, 3, 3, generator=g)
    want = F.conv3d(F.pad(x, (1, 1, 1, 1, 2, 0)).unsqueeze(0), w)[0]                        # (10, T, H, W)
    x8 = F.pad(x, (0, 0, 0, 0, 0, 0, 0, 5))
    got = (_im2col(x8, (3, 3, 3), (2, 1, 1)) @ conv_weight_2d(w).t()).reshape(4, 6, 7, 10).permute(3, 0, 1, 2)
    assert torch.allclose(got, want, atol=1e-4)
    # Conv2d weights become kt = 1
    w2 = torch.randn(5, 8, 3, 3, generator=g)
    x2 = torch.randn(8, 2, 5, 6, generator=g)
    want2 = F.conv3d(F.pad(x2, (1, 1, 1, 1)).unsqueeze(0), w2.unsqueeze(2))[0]
    got2 = (_im2col(x2, (1, 3, 3), (0, 1, 1)) @ conv_weight_2d(w2).t()).reshape(2, 5, 6, 5).permute(3, 0, 1, 2)
    assert torch.allclose(got2, want2, atol=1e-4)
    # folded input convolution: rows of W + 2 positions x 8 channels with zero borders, 64-element windows, taps (dt, dh)
    T, H, W = 4, 6, 7
    padded = torch.zeros(T, H, W + 2, 8)
    padded[:, :, 1:-1, :3] = x.permute(1, 2, 3, 0)
    flat = torch.cat([padded.reshape(-1), torch.zeros(64)])
    windows = torch.as_strided(flat, (T, H, W + 2, 64), (H * (W + 2) * 8, (W + 2) * 8, 8, 1))   # overlapping, pitch 8
    wf = fold_input_conv_weight(w)
    assert wf.shape == (10, 9 * 64)
    rows = _im2col(windows.permute(3, 0, 1, 2), (3, 3, 1), (2, 1, 0))                            # [T*H*(W+2), 9*64]
    gotf = (rows @ wf.t()).reshape(T, H, W + 2, 10)[:, :, :W].permute(3, 0, 1, 2)
    assert torch.allclose(gotf, want, atol=1e-4)
    # head: (3,1,1) convolution to tap-channels, then the nine shifted partial sums
    wh_src = torch.randn(3, 12, 3, 3, 3, generator=g)
    xh = torch.randn(12, 4, 5, 6, generator=g)
    wanth = F.conv3d(F.pad(xh, (1, 1, 1, 1, 2, 0)).unsqueeze(0), wh_src)[0]
    wt = head_tap_weight(wh_src)
    assert wt.shape == (36, 36) and float(wt.reshape(9, 4, -1)[:, 3].abs().max()) == 0.0
    part = (_im2col(xh, (3, 1, 1), (2, 0, 0)) @ wt.t()).reshape(4, 5, 6, 36).permute(3, 0, 1, 2)     # (36, T, H, W)
    pp = F.pad(part, (1, 1, 1, 1))
    goth = sum(pp[(dh * 3 + dw) * 4:(dh * 3 + dw) * 4 + 3, :, dh:dh + 5, dw:dw + 6] for dh in range(3) for dw in range(3))
    assert torch.allclose(goth, wanth, atol=1e-4)


def test_tile_plan_of_the_product_matches_the_oracle():
    from goal_force_b200.wan_vae import WanVideoVAEB200
    for (H, W, size, stride) in ((60, 104, (30, 52), (15, 26)), (90, 160, (30, 52), (15, 26)), (7, 9, (4, 5), (3, 3)),
                                 (34, 34, (34, 34), (18, 16)), (480, 832, (240, 416), (120, 208))):
        assert WanVideoVAEB200._tasks(H, W, size, stride) == V.tile_tasks(H, W, size, stride)


def _tiles_worker(rank, world, port, ret):
    import os
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from goal_force_b200.wan_vae import WanVideoVAEB200
        vae = object.__new__(WanVideoVAEB200)              # host logic only: no CUDA library, no weights
        vae.device = torch.device("cpu")
        tasks = WanVideoVAEB200._tasks(7, 9, (4, 5), (3, 3))
        computed = []

        def fn(task):                                      # a tile whose content identifies its task and its owner
            computed.append(task)
            h, h_, w, w_ = task
            return torch.full((3, 2, min(h_, 7) - h, min(w_, 9) - w), float(8 * h + w), dtype=torch.bfloat16)   # bf16-exact tags

        def shape_of(task):
            h, h_, w, w_ = task
            return (3, 2, min(h_, 7) - h, min(w_, 9) - w)

        seen = [(task, float(tile.float().mean()), tuple(tile.shape))
                for task, tile in vae._tiles(tasks, fn, shape_of, dist.group.WORLD)]
        assert [s_[0] for s_ in seen] == tasks                                   # the reference's task order, on every rank
        assert all(v == 8 * task[0] + task[2] and shp == shape_of(task) for task, v, shp in seen)
        assert computed == [task for i, task in enumerate(tasks) if i % world == rank]   # round-robin ownership
        single = [(task, float(tile.float().mean())) for task, tile in vae._tiles(tasks, fn, shape_of, None)]
        assert [(a, b) for a, b, _ in seen] == single
        if rank == 0:
            ret["n"] = len(tasks)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_tile_sharding_over_a_process_group_world2():
    """WanVideoVAEB200._tiles under gloo, world size 2: tiles are computed round-robin, broadcast from their owners and
    yielded in the reference's task order on every rank (the GPU test checks the decoded result bit for bit)."""
    import socket
    import torch.multiprocessing as mp
    with socket.socket() as s_:
        s_.bind(("127.0.0.1", 0))
        port = s_.getsockname()[1]
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_tiles_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret["n"] == 6
