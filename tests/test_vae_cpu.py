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
