"""Host side of the VideoModel variants (no GPU): for every supported constructor combination the product model has the
state_dict of the reference -- same keys, same order, same values under the same seed (the oracle's init_params is
pinned to the live reference's state_dict in tests/test_oracle_vs_reference.py) -- so reference checkpoints load."""
import pytest
import torch

from oracle import ta3n_oracle as orc
from ta3n_b200.models import VideoModel

VARIANTS = [dict(), dict(use_attn="none"), dict(use_attn="general"), dict(use_attn_frame="TransAttn"), dict(ens_DA="MCD"),
            dict(frame_aggregation="avgpool"), dict(frame_aggregation="avgpool", use_attn="none", ens_DA="MCD")]
N_PATH = {"trn-m": lambda R: 6 + 6 * R + 6, "avgpool": lambda R: 12}


@pytest.mark.parametrize("kw", VARIANTS, ids=lambda kw: "-".join(f"{k}={v}" for k, v in kw.items()) or "default")
def test_state_dict_equals_reference_layout_and_init(kw):
    agg = kw.get("frame_aggregation", "trn-m")
    cfg = orc.PathConfig(num_class=7, num_segments=4, fc_dim=128, use_attn=kw.get("use_attn", "TransAttn"),
                         use_attn_frame=kw.get("use_attn_frame", "none"), ens_DA=kw.get("ens_DA", "none"),
                         frame_aggregation=agg)
    want = orc.init_params(cfg, seed=3)
    torch.manual_seed(3)
    m = VideoModel(7, "video", agg, "RGB", train_segments=4, val_segments=4, fc_dim=128, verbose=False,
                   use_attn=cfg.use_attn, use_attn_frame=cfg.use_attn_frame, ens_DA=cfg.ens_DA)
    sd = m.state_dict()
    assert list(sd.keys()) == list(want.keys())
    for k in sd:
        assert torch.equal(sd[k], want[k]), k
    n = N_PATH[agg](3) + (4 if cfg.use_attn == "general" else 0)
    pp = m.path_parameters()
    assert len(pp) == n and len({id(t) for t in pp}) == n
    used = set(orc.used_param_names(want)) - {k for k in want if k.startswith("fc_classifier_video_source_2")}
    names = {id(v): k for k, v in m.named_parameters()}
    assert {names[id(t)] for t in pp} == used      # exactly the parameters the reference gives a gradient on this path
