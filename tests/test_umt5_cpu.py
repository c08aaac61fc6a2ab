"""umT5 prompt encoder, CPU side: the oracle restatement against the vectors the reference's own
WanTextEncoder produced (tests/golden/umt5.pt, oracle/gen_golden.py::gen_umt5), and the host-side bucket table."""
import pytest
import torch

from goal_force_b200.umt5 import UMT5Config, relative_position_bucket
from oracle import umt5_oracle as U
from oracle import wan_dit_oracle as O


def _case(golden_dir):
    g = torch.load(golden_dir / "umt5.pt", weights_only=False)["tiny"]
    c = g["cfg"]
    sd = U.random_state_dict(seed=g["weight_seed"], **c)
    ids, mask = U.synthetic_prompt(c["vocab"], g["batch"], g["L"], g["valid"], seed=g["prompt_seed"])
    return g, c, sd, ids, mask


def test_oracle_matches_reference_vectors(golden_dir):
    g, c, sd, ids, mask = _case(golden_dir)
    kw = dict(num_heads=c["num_heads"], num_layers=c["num_layers"], num_buckets=c["num_buckets"])
    with torch.no_grad():
        assert torch.equal(U.encoder(sd, ids, mask, **kw), g["out_fp32"])
        sdb = {k: v.to(torch.bfloat16) for k, v in sd.items()}
        out_bf = U.encoder(sdb, ids, mask, **kw)
    assert out_bf.dtype == torch.bfloat16
    assert O.rel_l2(out_bf, g["out_bf16"]) < 1e-6          # same ops in the same order (thread count may regroup sums)
    # the padding mask matters: without it the padded sequence changes
    with torch.no_grad():
        nomask = U.encoder(sd, ids, None, **kw)
    assert torch.equal(nomask[0], g["out_fp32"][0]) or O.rel_l2(nomask[0], g["out_fp32"][0]) < 1e-6
    assert O.rel_l2(nomask[1], g["out_fp32"][1]) > 1e-3


def test_bucket_table_matches_reference(golden_dir):
    g, c, *_ = _case(golden_dir)
    L = g["L"]
    t = relative_position_bucket(L, L, c["num_buckets"])
    assert t.dtype == torch.int32 and t.shape == (2 * L - 1,)
    i = torch.arange(L).unsqueeze(1)
    j = torch.arange(L).unsqueeze(0)
    assert torch.equal(t[(j - i + L - 1)], g["buckets"])
    # umT5-XXL defaults: 32 buckets, 16 per direction, exact up to 8, log-spaced up to 128
    t512 = relative_position_bucket(512, 512, 32)
    assert int(t512.min()) == 0 and int(t512.max()) == 31 and int(t512[511]) == 0
    assert t512[511 + 3] == 16 + 3 and t512[511 - 3] == 3 and t512[511 + 400] == 31 and t512[511 - 400] == 15


def test_config_defaults_are_umt5_xxl():
    c = UMT5Config()
    assert (c.vocab, c.dim, c.dim_attn, c.dim_ffn, c.num_heads, c.num_layers, c.num_buckets) == \
        (256384, 4096, 4096, 10240, 64, 24, 32)


@pytest.mark.reference
def test_state_dict_keys_match_live_reference():
    from oracle import ref_shim
    te = ref_shim.load_module("diffsynth.models.wan_video_text_encoder")
    c = dict(vocab=97, dim=256, dim_attn=256, dim_ffn=512, num_heads=4, num_layers=2, num_buckets=32)
    m = te.WanTextEncoder(shared_pos=False, **c)
    assert sorted(m.state_dict()) == sorted(U.random_state_dict(**c))
