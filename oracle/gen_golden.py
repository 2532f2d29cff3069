"""Generate ``tests/golden/ta3n_golden.npz`` by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python -m oracle.gen_golden

Every case fixes: model seed (torch.manual_seed before constructing the reference
VideoModel), input seed, config, beta/gamma.  Inputs and parameters are NOT
stored (they regenerate from the seeds on the same image); their checksums are,
so drift is detected rather than silently compared against.
Stored per case: loss, all small outputs in full, feat_fc checksums + a strided
sample, the L2 norm of every parameter gradient and strided samples of each.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import ref_shims  # noqa: E402
from oracle import ta3n_oracle as orc  # noqa: E402

GOLDEN_PATH = os.path.join(os.path.dirname(HERE), "tests", "golden", "ta3n_golden.npz")
SAMPLE_STRIDE = 1009     # prime stride for samples of large tensors

CASES = {
    # name: dict(B_s, B_t, T, C, F, train(masks), use_attn, use_attn_frame, model_seed, input_seed)
    "fingerprint_c12": dict(bs=8, bt=8, T=5, C=12, F=512, train=False, use_attn="TransAttn", attn_frame="none"),
    "cfg1_small_c5": dict(bs=8, bt=8, T=5, C=5, F=512, train=False, use_attn="TransAttn", attn_frame="none"),
    "cfg1_train_masked": dict(bs=8, bt=8, T=5, C=5, F=512, train=True, use_attn="TransAttn", attn_frame="none"),
    "ragged_6_3": dict(bs=6, bt=3, T=5, C=12, F=512, train=True, use_attn="TransAttn", attn_frame="none"),
    "t9_attnframe": dict(bs=4, bt=4, T=9, C=12, F=512, train=True, use_attn="TransAttn", attn_frame="TransAttn"),
    "noattn_f256": dict(bs=5, bt=5, T=5, C=12, F=256, train=False, use_attn="none", attn_frame="none"),
    "t3_f2048": dict(bs=2, bt=2, T=3, C=6, F=2048, train=True, use_attn="TransAttn", attn_frame="none"),
    "general_attn": dict(bs=7, bt=5, T=5, C=12, F=512, train=True, use_attn="general", attn_frame="none"),
    # frame_aggregation='avgpool' (key "agg"; default 'trn-m'): the paper's baseline aggregation, with / without attention
    "avgpool_transattn": dict(bs=6, bt=5, T=5, C=12, F=512, train=True, use_attn="TransAttn", attn_frame="none",
                              agg="avgpool"),
    "avgpool_noattn_f256": dict(bs=4, bt=7, T=3, C=7, F=256, train=False, use_attn="none", attn_frame="none",
                                agg="avgpool"),
}
BETA = (0.75, 0.75, 0.5)
GAMMA = 0.003
MODEL_SEED = 1234
INPUT_SEED = 4321
MASK_SEED = 777
DROPOUT = 0.5


def case_config(c) -> orc.PathConfig:
    return orc.PathConfig(num_class=c["C"], num_segments=c["T"], fc_dim=c["F"], dropout_i=DROPOUT,
                          dropout_v=DROPOUT, use_attn=c["use_attn"], use_attn_frame=c["attn_frame"],
                          frame_aggregation=c.get("agg", "trn-m"))


def case_inputs(c):
    """Inputs / labels / keep-masks for a case -- shared by the generator and the tests."""
    cfg = case_config(c)
    g = torch.Generator().manual_seed(INPUT_SEED)
    xs = torch.randn(c["bs"], c["T"], orc.FEATURE_DIM, generator=g)
    xt = torch.randn(c["bt"], c["T"], orc.FEATURE_DIM, generator=g)
    labels = torch.arange(c["bs"]) % c["C"]
    masks = None
    if c["train"]:
        gm = torch.Generator().manual_seed(MASK_SEED)
        keep = 1.0 - DROPOUT
        masks = {
            "i_source": (torch.rand(c["bs"] * c["T"], cfg.shared_dim, generator=gm) < keep).to(torch.uint8),
            "i_target": (torch.rand(c["bt"] * c["T"], cfg.shared_dim, generator=gm) < keep).to(torch.uint8),
            "v_source": (torch.rand(c["bs"], cfg.video_dim, generator=gm) < keep).to(torch.uint8),
            "v_target": (torch.rand(c["bt"], cfg.video_dim, generator=gm) < keep).to(torch.uint8),
        }
    return cfg, xs, xt, labels, masks


def sample(t: torch.Tensor) -> np.ndarray:
    flat = t.detach().reshape(-1).double()
    return flat[::SAMPLE_STRIDE].numpy().copy()


def checksum(t: torch.Tensor) -> np.ndarray:
    d = t.detach().double()
    return np.array([d.sum().item(), d.norm().item()])


def reference_loss(outs, labels, use_attn="TransAttn"):
    """Loss composition exactly as main.py:446, 508-538, 559-562 (uSv + RevGrad + attentive_entropy), on the
    reference's own 10-tuple, using the reference's loss.py."""
    _, _, ref_loss = ref_shims.load()
    (attn_s, out_s, _, pd_s, feat_s, attn_t, out_t, _, pd_t, feat_t) = outs
    ce = torch.nn.CrossEntropyLoss()
    loss = ce(out_s, labels)
    pred_domain_all = []
    for l in range(3):
        ps = pd_s[l].view(-1, pd_s[l].size()[-1])
        pt = pd_t[l].view(-1, pd_t[l].size()[-1])
        dom = torch.cat((torch.zeros(ps.size(0)).long(), torch.ones(pt.size(0)).long()), 0)
        pred = torch.cat((ps, pt), 0)
        pred_domain_all.append(pred)
        loss = loss + ce(pred, dom)
    if use_attn != "none":
        loss = loss + GAMMA * ref_loss.attentive_entropy(torch.cat((out_s, out_t), 0), pred_domain_all[1])
    return loss


def run_reference(c, dtype=torch.float32):
    ref_models, _, ref_loss = ref_shims.load()
    cfg, xs, xt, labels, masks = case_inputs(c)
    xs, xt = xs.to(dtype), xt.to(dtype)
    torch.manual_seed(MODEL_SEED)
    model = ref_models.VideoModel(c["C"], "video", c.get("agg", "trn-m"), "RGB", train_segments=c["T"], val_segments=c["T"],
                                  add_fc=1, fc_dim=c["F"], dropout_i=DROPOUT, dropout_v=DROPOUT,
                                  partial_bn=False, use_bn="none", ens_DA="none", use_attn=c["use_attn"],
                                  n_attn=1, use_attn_frame=c["attn_frame"], share_params="Y", verbose=False)
    model = model.to(dtype)
    if c["train"]:
        model.train()
        model.dropout_i = ref_shims.InjectedDropout(DROPOUT, [masks["i_source"], masks["i_target"]])
        model.dropout_v = ref_shims.InjectedDropout(DROPOUT, [masks["v_source"], masks["v_target"]])
    else:
        model.eval()
    outs = model(xs, xt, list(BETA), 0, is_train=True, reverse=False)
    loss = reference_loss(outs, labels, c["use_attn"])
    loss.backward()
    return model, outs, loss, (xs, xt)


def main():
    blob = {}
    meta = {"beta": BETA, "gamma": GAMMA, "model_seed": MODEL_SEED, "input_seed": INPUT_SEED,
            "mask_seed": MASK_SEED, "dropout": DROPOUT, "stride": SAMPLE_STRIDE, "cases": CASES,
            "torch": torch.__version__}
    for name, c in CASES.items():
        model, outs, loss, (xs, xt) = run_reference(c)
        # the same reference in float64: |fp32 - fp64| is the reference's own rounding noise, stored
        # per tensor so that tests can tell cancellation noise from real disagreement
        model64, outs64, loss64, _ = run_reference(c, torch.float64)
        grads64 = {n: p.grad for n, p in model64.named_parameters()}
        (attn_s, out_s, _, pd_s, feat_s, attn_t, out_t, _, pd_t, feat_t) = outs
        k = name + "/"
        blob[k + "loss"] = np.array(loss.item())
        blob[k + "noise/loss"] = np.array(abs(loss.item() - loss64.item()))
        blob[k + "in_checksum"] = np.concatenate([checksum(xs), checksum(xt)])
        for dom, attn, out, pd, feat in (("s", attn_s, out_s, pd_s, feat_s), ("t", attn_t, out_t, pd_t, feat_t)):
            blob[k + f"attn_{dom}"] = attn.detach().numpy()
            blob[k + f"out_{dom}"] = out.detach().numpy()
            blob[k + f"pred_rel_{dom}"] = pd[0].detach().numpy()
            blob[k + f"pred_video_{dom}"] = pd[1].detach().numpy()
            blob[k + f"pred_frame_{dom}"] = pd[2].detach().numpy()
            blob[k + f"feat_video_{dom}"] = feat[1].detach().numpy()
            blob[k + f"feat_fc_{dom}_checksum"] = checksum(feat[2])
            blob[k + f"feat_fc_{dom}_sample"] = sample(feat[2])
        used = []
        for pname, prm in model.named_parameters():
            blob[k + "param_checksum/" + pname] = checksum(prm)
            if prm.grad is None:
                continue
            used.append(pname)
            blob[k + "grad_norm/" + pname] = np.array(prm.grad.double().norm().item())
            blob[k + "grad_sample/" + pname] = sample(prm.grad)
            blob[k + "grad_noise/" + pname] = np.array((prm.grad.double() - grads64[pname]).norm().item())
            blob[k + "grad_sample_noise/" + pname] = np.array(
                np.linalg.norm(sample(prm.grad) - sample(grads64[pname])))
        meta.setdefault("used_params", {})[name] = used
        print(f"{name}: loss={loss.item():.8f} used_params={len(used)}")
    blob["meta_json"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    os.makedirs(os.path.dirname(GOLDEN_PATH), exist_ok=True)
    np.savez_compressed(GOLDEN_PATH, **blob)
    print("wrote", GOLDEN_PATH, os.path.getsize(GOLDEN_PATH), "bytes")


if __name__ == "__main__":
    main()
