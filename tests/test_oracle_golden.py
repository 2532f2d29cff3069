"""The CPU oracle restatement vs the committed golden vectors (made by the live reference)."""
import pytest
import torch

from oracle import gen_golden
from oracle import ta3n_oracle as orc
from tests.golden_util import (TOL_FP32, assert_close, check_grads_against_golden,
                               check_outputs_against_golden, load_golden)

Z, META = load_golden()


def test_relation_tuples_match_survey_appendix_a():
    t5 = orc.relation_tuples(5)
    assert t5 == [[(0, 1, 2, 3, 4)],
                  [(0, 1, 2, 3), (0, 1, 3, 4), (1, 2, 3, 4)],
                  [(0, 1, 2), (0, 2, 4), (1, 2, 4)],
                  [(0, 1), (1, 2), (2, 3)]]
    t9 = orc.relation_tuples(9)
    assert sum(len(r) for r in t9) == 22
    assert sum(len(t) for r in t9 for t in r) == 114
    assert t9[7] == [(0, 1), (1, 6), (3, 7)]
    assert t9[4] == [(0, 1, 2, 3, 4), (0, 2, 3, 6, 7), (1, 2, 4, 6, 8)]
    assert orc.relation_tuples(2) == [[(0, 1)]]


@pytest.mark.parametrize("case", list(gen_golden.CASES))
def test_oracle_reproduces_reference_golden(case):
    c = gen_golden.CASES[case]
    cfg, xs, xt, labels, masks = gen_golden.case_inputs(c)
    params = orc.init_params(cfg, seed=META["model_seed"])
    # init parity: same seed -> same parameter values as the reference's constructor
    for name, t in params.items():
        key = f"{case}/param_checksum/{name}"
        if key in Z.files:
            cs = Z[key]
            d = t.double()
            assert abs(d.sum().item() - cs[0]) <= 1e-6 * max(1.0, abs(cs[0])), name
            assert abs(d.norm().item() - cs[1]) <= 1e-6 * max(1.0, cs[1]), name
    ics = Z[f"{case}/in_checksum"]
    assert abs(xs.double().sum().item() - ics[0]) < 1e-6 * max(1, abs(ics[0]))
    loss, outs, grads = orc.train_step(params, xs, xt, labels, META["beta"], cfg, META["gamma"],
                                       train=c["train"], masks=masks)
    assert_close(loss, Z[f"{case}/loss"], TOL_FP32, "loss")
    check_outputs_against_golden(Z, case, outs, TOL_FP32, META["stride"])
    used = META["used_params"][case]
    assert sorted(used) == sorted(orc.used_param_names(params))
    check_grads_against_golden(Z, case, grads, used, 2e-4, META["stride"])


def test_fingerprint_matches_survey_probe():
    # loss = 4.57745266 for the C=12 fingerprint config
    assert abs(float(Z["fingerprint_c12/loss"]) - 4.57745266) < 5e-7


@pytest.mark.parametrize("attn_frame", ["none", "TransAttn"])
def test_oracle_gates_reproduce_the_plain_run(attn_frame):
    """Pinning the ReLU pattern to the pattern of the plain run must change nothing (gate plumbing)."""
    cfg = orc.PathConfig(num_class=7, num_segments=4, fc_dim=64, dropout_i=0.0, dropout_v=0.0,
                         use_attn_frame=attn_frame)
    params = {k: (v.double() if v.dtype.is_floating_point else v) for k, v in orc.init_params(cfg, seed=3).items()}
    xs, xt, labels = orc.synthetic_batch(6, cfg, dtype=torch.float64)
    xt = xt[:4]
    beta = (0.75, 0.75, 0.5)
    l0, _, g0 = orc.train_step(params, xs, xt, labels, beta, cfg, 0.003, train=True)
    gates = orc.activation_pattern(params, xs, xt, beta, cfg)
    l1, _, g1 = orc.train_step(params, xs, xt, labels, beta, cfg, 0.003, train=True, gates=gates)
    assert abs(l0.item() - l1.item()) < 1e-12
    for k in g0:
        assert torch.allclose(g0[k], g1[k], rtol=1e-10, atol=1e-14), k
    # flipping a handful of units changes gradients by far more than rounding: the sqrt(eps) effect
    flipped = {k: ([t.clone() for t in v] if isinstance(v, list) else v.clone()) for k, v in gates.items()}
    flipped["frame_disc"][0, :8] = ~flipped["frame_disc"][0, :8]
    _, _, g2 = orc.train_step(params, xs, xt, labels, beta, cfg, 0.003, train=True, gates=flipped)
    k = "fc_feature_domain.weight"
    assert (g2[k] - g0[k]).norm() / g0[k].norm() > 1e-3
