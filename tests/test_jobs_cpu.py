"""Batched job driver (goal_force_b200.jobs): CSV sharding, Direct Force rows, host-side control-video prefetch, and
the replica x cfg x sp layout -- everything that runs without a GPU.  The denoiser is replaced by a recording stub."""
import json
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from goal_force_b200 import jobs as J
from goal_force_b200.pipeline import ParallelLayout
from oracle import control_channels_oracle as CC


def test_shard_contiguous_matches_reference_semantics():
    items = list("abcde")
    assert J.shard_contiguous(items, 2, 0) == ["a", "b", "c"] and J.shard_contiguous(items, 2, 1) == ["d", "e"]
    for n in range(0, 13):
        for w in range(1, 9):
            parts = [J.shard_contiguous(list(range(n)), w, d) for d in range(w)]
            assert sum(parts, []) == list(range(n))
            sizes = [len(p) for p in parts]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)


@pytest.mark.reference
def test_shard_contiguous_equals_reference_function():
    import importlib.util
    from oracle import ref_shim
    spec = importlib.util.spec_from_file_location("ref_inf_utils", os.path.join(ref_shim.REFERENCE_ROOT,
                                                                                "scripts/inference/utils.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    for n in range(0, 11):
        for w in range(1, 7):
            for d in range(w):
                assert J.shard_contiguous(list(range(n)), w, d) == mod.split_list_across_devices_contiguous(list(range(n)), w, d)


@pytest.mark.reference
def test_read_rows_equals_pandas_rows():
    import glob
    import pandas
    from oracle import ref_shim
    csvs = sorted(glob.glob(os.path.join(ref_shim.REFERENCE_ROOT, "datasets/examples/*/*.csv")))
    csvs = [c for c in csvs if "canny" not in c]
    assert len(csvs) == 12
    for c in csvs[:4]:
        mine = J.read_rows(c)[0]
        ref = pandas.read_csv(c).iloc[0].to_dict()
        for k in J._NUMERIC:
            assert float(mine[k]) == float(ref[k]), (c, k)
        np.random.seed(0)
        from goal_force_b200 import control_channels as P
        a = P.control_video_from_csv_row(mine)
        np.random.seed(0)
        b = P.control_video_from_csv_row(ref)
        assert torch.equal(a, b)


def test_direct_force_rows_through_prefetcher_match_reference_digests(golden_dir):
    """BASELINE configs[3]: Direct Force rows (projectile force + mass channels).  The prefetcher's per-job
    RandomState(0) consumes the generator exactly like np.random.seed(0) + the reference dataset did when
    oracle/gen_golden.py produced the digests."""
    g = json.loads((golden_dir / "control_channels.json").read_text())
    names = ["_pendulum", "_toycar", "_cantaloupes", "_paw_tool2"]
    jobs = [J.Job(row=J.direct_force_row(g["rows"][n], 250.0, 37.0, 2.5), index=i) for i, n in enumerate(names)]
    pre = J.ControlVideoPrefetcher(jobs, 81, 480, 832, rng_seed=0)
    pre.start(0)
    for k, n in enumerate(names):
        v = pre.get(k)
        pre.start(k + 1)
        assert CC.digest(v) == g["direct_force"][n], n
        assert float(v[..., 1].abs().max()) == 0.0 and float(v[..., 0].max()) > 0.9 and float(v[..., 2].max()) > 0.9


class _StubDenoiser:
    def __init__(self):
        self.calls = []

    def __call__(self, noise, ctx_p, ctx_n, y=None, control_latents=None, **kw):
        self.calls.append((tuple(noise.shape), float(control_latents.float().abs().sum()), kw["num_inference_steps"]))
        return noise + control_latents.mean()


def test_batch_driver_runs_rows_in_order_with_prefetch(golden_dir):
    g = json.loads((golden_dir / "control_channels.json").read_text())
    rows = [J.direct_force_row(g["rows"][n], 100.0 + 50 * i, 10.0 * i, 1.5 + i) for i, n in
            enumerate(["_golf", "_tennis", "_soccer_tool"])]
    den = _StubDenoiser()
    cond = lambda row: dict(context_posi=torch.zeros(1, 4, 8), context_nega=torch.zeros(1, 4, 8), y=None)  # noqa: E731
    drv = J.BatchDriver(den, J.synthetic_control_encoder("cpu"), cond, num_frames=9, height=480, width=832,
                        num_inference_steps=3, device="cpu")
    done = []
    out = drv.run(rows, seed=5, on_done=lambda job: done.append(job.index))
    assert done == [0, 1, 2] and [j.index for j in out] == [0, 1, 2]
    assert all(c[0] == (1, 16, 3, 60, 104) and c[2] == 3 for c in den.calls)
    assert len({c[1] for c in den.calls}) == 3                         # three different control videos reached the denoiser
    assert all(j.info["control_digest_channels"] == [True, False, True] for j in out)      # direct force: ch0 + ch2
    lat = J.synthetic_control_encoder("cpu")(out[0].control_video)
    assert lat.shape == (1, 16, 3, 60, 104) and lat.dtype == torch.bfloat16


def test_layout_with_replicas():
    lay = ParallelLayout(world_size=8, rank=5, cfg_size=2, replicas=4)       # 4 replicas x cfg 2 x sp 1
    assert (lay.group_size, lay.sp_size, lay.replica, lay.cfg_index, lay.sp_index) == (2, 1, 2, 1, 0)
    assert lay.cfg_ranks() == [4, 5] and lay.sp_ranks() == [5]
    lay = ParallelLayout(world_size=8, rank=6, cfg_size=2, replicas=1)       # cfg 2 x sp 4 (BASELINE configs[3])
    assert (lay.sp_size, lay.cfg_index, lay.sp_index) == (4, 1, 2)
    assert lay.sp_ranks() == [4, 5, 6, 7] and lay.cfg_ranks() == [2, 6]
    lay = ParallelLayout(world_size=8, rank=3, cfg_size=1, replicas=2)       # 2 replicas x sp 4
    assert (lay.replica, lay.sp_ranks(), lay.cfg_ranks()) == (0, [0, 1, 2, 3], [3])
    with pytest.raises(ValueError):
        ParallelLayout(world_size=8, rank=0, cfg_size=2, replicas=3)


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ret):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from goal_force_b200.pipeline import ParallelContext
        par = ParallelContext(ParallelLayout(world_size=world, rank=rank, cfg_size=1, replicas=2))
        assert par.sp is None and par.cfg_group is None
        den = _StubDenoiser()
        cond = lambda row: dict(context_posi=torch.zeros(1, 4, 8), context_nega=None, y=None)  # noqa: E731
        drv = J.BatchDriver(den, J.synthetic_control_encoder("cpu"), cond, parallel=par, num_frames=5, height=32,
                            width=48, num_inference_steps=2, cfg_scale=1.0, device="cpu")
        rows = [dict(projectile_force_magnitude=100.0 + i, projectile_force_angle=5.0 * i, projectile_coordx=10 + i,
                     projectile_coordy=12, projectile_mass=2.0, target_indirect_force_magnitude=-1.0,
                     target_indirect_force_angle=0.0, target_coordx=30, target_coordy=20, target_mass=-1.0,
                     width=48, height=32) for i in range(5)]
        out = drv.run(rows)
        got = [None] * world
        dist.all_gather_object(got, [j.index for j in out])
        if rank == 0:
            ret["idx"] = got
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_replica_groups_take_contiguous_shards_world2():
    """world_size 2 (gloo): two replicas, five rows -> replica 0 takes rows 0-2, replica 1 rows 3-4; no data-path
    collective is involved (the path shards over independent rows)."""
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), ret), nprocs=2, join=True)
    assert ret["idx"] == [[0, 1, 2], [3, 4]]


def test_decode_stage_and_uint8_frames():
    """The optional decode stage runs after every row's denoising (wan_video_new.py:731-734), and video_to_uint8 is the
    reference's vae_output_to_video arithmetic (diffsynth/utils/__init__.py:76-91)."""
    den = _StubDenoiser()
    cond = lambda row: dict(context_posi=torch.zeros(1, 4, 8), context_nega=None, y=None)  # noqa: E731
    seen = []

    def decode(lat):
        seen.append(tuple(lat.shape))
        return torch.linspace(-1.2, 1.2, 3 * 5 * 4 * 6).reshape(1, 3, 5, 4, 6)

    drv = J.BatchDriver(den, J.synthetic_control_encoder("cpu"), cond, num_frames=5, height=32, width=48,
                        num_inference_steps=2, cfg_scale=1.0, device="cpu", decode=decode)
    rows = [dict(projectile_force_magnitude=100.0, projectile_force_angle=5.0, projectile_coordx=10, projectile_coordy=12,
                 projectile_mass=2.0, target_indirect_force_magnitude=-1.0, target_indirect_force_angle=0.0,
                 target_coordx=30, target_coordy=20, target_mass=-1.0, width=48, height=32) for _ in range(2)]
    out = drv.run(rows)
    assert len(seen) == 2 and all(j.video is not None and j.video.shape == (1, 3, 5, 4, 6) for j in out)
    frames = J.video_to_uint8(out[0].video)
    assert frames.shape == (5, 4, 6, 3) and frames.dtype == torch.uint8
    v = out[0].video[0].permute(1, 2, 3, 0)
    want = ((v - (-1)) * (255 / (1 - (-1)))).clip(0, 255).to(dtype=torch.uint8)        # the reference's expression
    assert torch.equal(frames, want)
    assert int(frames.min()) == 0 and int(frames.max()) == 255
