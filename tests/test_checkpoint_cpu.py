"""safetensors ingestion (host logic, no GPU): shard merging, lazy reads, the 'pipe.controlnet.' checkpoint prefix."""
import pytest
import torch
from safetensors.torch import save_file

from goal_force_b200.checkpoint import SafetensorsStateDict
from goal_force_b200.wan_dit import _PrefixTolerant


def test_sharded_state_dict_round_trip(tmp_path):
    g = torch.Generator().manual_seed(0)
    a = {"blocks.0.self_attn.q.weight": torch.randn(8, 8, generator=g).bfloat16(), "patch_embedding.bias": torch.randn(8, generator=g)}
    b = {"blocks.1.ffn.0.weight": torch.randn(16, 8, generator=g).bfloat16(), "head.modulation": torch.randn(1, 2, 8, generator=g)}
    save_file(a, str(tmp_path / "model-00001-of-00002.safetensors"))
    save_file(b, str(tmp_path / "model-00002-of-00002.safetensors"))
    sd = SafetensorsStateDict(sorted(tmp_path.glob("*.safetensors")))
    assert len(sd) == 4 and "head.modulation" in sd and "nope" not in sd
    for k, v in {**a, **b}.items():
        got = sd[k]
        assert got.dtype == v.dtype and torch.equal(got, v)
    with pytest.raises(KeyError):
        sd["missing.key"]
    with pytest.raises(FileNotFoundError):
        SafetensorsStateDict(tmp_path / "absent.safetensors")


def test_controlnet_checkpoint_prefix(tmp_path):
    w = torch.randn(4, 4)
    save_file({"pipe.controlnet.controlnet_zero_convs_after.0.bias": w}, str(tmp_path / "step-3000.safetensors"))
    sd = _PrefixTolerant(SafetensorsStateDict(tmp_path / "step-3000.safetensors"), "pipe.controlnet.")
    assert torch.equal(sd["controlnet_zero_convs_after.0.bias"], w)       # as load_controlnet_weights strips it
    with pytest.raises(KeyError):
        sd["controlnet_zero_convs_after.1.bias"]


def test_vae_checkpoint_forms(tmp_path):
    """read_vae_state_dict accepts the forms the reference's loader accepts: plain pickle, {'model_state': ...},
    'model.'-prefixed keys, safetensors; anything else is refused."""
    import pytest
    from safetensors.torch import save_file
    from goal_force_b200.checkpoint import read_vae_state_dict
    from oracle import wan_vae_oracle as V
    sd = V.random_state_dict(dim=32, seed=0)
    torch.save(sd, tmp_path / "plain.pth")
    torch.save({"model_state": sd}, tmp_path / "wrapped.pth")
    torch.save({"model." + k: v for k, v in sd.items()}, tmp_path / "prefixed.pth")
    save_file({k: v.contiguous() for k, v in sd.items()}, str(tmp_path / "vae.safetensors"))
    for name in ("plain.pth", "wrapped.pth", "prefixed.pth", "vae.safetensors"):
        got = read_vae_state_dict(tmp_path / name)
        assert sorted(got) == sorted(sd), name
        assert torch.equal(got["decoder.head.2.weight"], sd["decoder.head.2.weight"])
    torch.save({"foo": torch.zeros(1)}, tmp_path / "other.pth")
    with pytest.raises(KeyError):
        read_vae_state_dict(tmp_path / "other.pth")
    with pytest.raises(FileNotFoundError):
        read_vae_state_dict(tmp_path / "missing.pth")
