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
