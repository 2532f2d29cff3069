"""Pin the oracle against the LIVE reference (build container only)."""
import pytest
import torch

from oracle import gen_golden, ref_shims
from oracle import ta3n_oracle as orc
from tests.golden_util import TOL_FP32, assert_close

pytestmark = pytest.mark.skipif(not ref_shims.available(), reason="/root/reference not present")


@pytest.mark.parametrize("case", ["cfg1_train_masked", "t9_attnframe", "noattn_f256"])
def test_oracle_equals_live_reference(case):
    c = gen_golden.CASES[case]
    model, outs_ref, loss_ref, _ = gen_golden.run_reference(c)
    cfg, xs, xt, labels, masks = gen_golden.case_inputs(c)
    params = {k: v.detach().clone() for k, v in model.state_dict().items()}
    loss, outs, grads = orc.train_step(params, xs, xt, labels, gen_golden.BETA, cfg, gen_golden.GAMMA,
                                       train=c["train"], masks=masks)
    assert_close(loss, loss_ref, TOL_FP32, "loss")
    flat_ref = [outs_ref[0], outs_ref[1], *outs_ref[3], *outs_ref[4], outs_ref[5], outs_ref[6], *outs_ref[8], *outs_ref[9]]
    flat = [outs[0], outs[1], *outs[3], *outs[4], outs[5], outs[6], *outs[8], *outs[9]]
    for i, (a, b) in enumerate(zip(flat, flat_ref)):
        assert a.shape == b.shape
        assert_close(a, b, TOL_FP32, f"output {i}")
    for name, prm in model.named_parameters():
        if prm.grad is None:
            assert name not in grads
        else:
            assert_close(grads[name], prm.grad, 2e-4, f"grad {name}")


def test_reference_state_dict_keys_match_oracle_init():
    ref_models, _, _ = ref_shims.load()
    torch.manual_seed(7)
    m = ref_models.VideoModel(12, "video", "trn-m", "RGB", train_segments=5, val_segments=5, add_fc=1,
                              fc_dim=512, partial_bn=False, use_bn="none", ens_DA="none",
                              use_attn="TransAttn", share_params="Y", verbose=False)
    p = orc.init_params(orc.PathConfig(num_class=12, num_segments=5, fc_dim=512), seed=7)
    sd = m.state_dict()
    assert list(sd.keys()) == list(p.keys())
    for k in sd:
        assert sd[k].shape == p[k].shape, k
        assert torch.equal(sd[k], p[k]), k
