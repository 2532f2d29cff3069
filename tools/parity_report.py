"""Print normwise relative errors (vs the fp64 CPU oracle) of every output and parameter gradient
of the CUDA path, per GEMM engine.  Run on a GPU box:  python tools/parity_report.py [B] [T]
X3=1 adds a column for the experimental 'tf32x3' engine."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ta3n_b200
from oracle import ta3n_oracle as orc
from ta3n_b200.models import VideoModel
from ta3n_b200.loss import ta3n_loss

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
T = int(sys.argv[2]) if len(sys.argv) > 2 else 5
dev = torch.device("cuda:0")
cfg = orc.PathConfig(num_class=12, num_segments=T, fc_dim=512, dropout_i=0.0, dropout_v=0.0)
params = orc.init_params(cfg, seed=1234)
if os.environ.get("PERTURB", "1") == "1":
    g = torch.Generator().manual_seed(5)
    for k in params:
        if params[k].dtype.is_floating_point and k.startswith(orc.USED_PARAM_PREFIXES) and "weight" in k:
            params[k] = params[k] + 0.02 * torch.randn(params[k].shape, generator=g)
xs, xt, labels = orc.synthetic_batch(B, cfg)
beta = (0.75, 0.75, 0.5)
p64 = {k: (v.double() if v.dtype.is_floating_point else v) for k, v in params.items()}
l64, o64, g64 = orc.train_step(p64, xs.double(), xt.double(), labels, beta, cfg, 0.003, train=True)
l32, o32, g32 = orc.train_step(params, xs, xt, labels, beta, cfg, 0.003, train=True)
flat = lambda o: [o[0], o[1], *o[3], *o[4], o[5], o[6], *o[8], *o[9]]
names = ["attn_s", "out_s", "pred_rel_s", "pred_vid_s", "pred_frame_s", "feat_pred_s", "feat_video_s", "feat_fc_s",
         "attn_t", "out_t", "pred_rel_t", "pred_vid_t", "pred_frame_t", "feat_pred_t", "feat_video_t", "feat_fc_t"]
def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-300)).item()
res = {}
X3 = os.environ.get("X3", "0") == "1"
for eng in ("fp32", "tf32") + (("tf32x3",) if X3 else ()):
    ta3n_b200.set_gemm_engine(eng)
    m = VideoModel(12, "video", "trn-m", "RGB", train_segments=T, val_segments=T, fc_dim=512, dropout_i=0.0,
                   dropout_v=0.0, partial_bn=False, verbose=False)
    m.load_state_dict(params); m = m.to(dev).train()
    outs = m(xs.to(dev), xt.to(dev), list(beta), 0, True, False)
    loss = ta3n_loss(outs, labels.to(dev), 0.003); loss.backward()
    r = {"loss": rel(loss, l64)}
    for n, a, b in zip(names, flat(outs), flat(o64)):
        r["out:" + n] = rel(a, b)
    for n, p in m.named_parameters():
        if n in g64:
            r["grad:" + n] = rel(p.grad, g64[n])
    res[eng] = r
# tf32 engine again, but against the fp64 oracle evaluated ON THE ACTIVATION PATTERN the CUDA forward realised
from ta3n_b200.train import TrainStep
ta3n_b200.set_gemm_engine("tf32")
m = VideoModel(12, "video", "trn-m", "RGB", train_segments=T, val_segments=T, fc_dim=512, dropout_i=0.0,
               dropout_v=0.0, partial_bn=False, verbose=False)
m.load_state_dict(params); m = m.to(dev).train()
step = TrainStep(m, B, B, beta, gamma=0.003, use_graph=False)
lossp = step(xs, xt, labels); torch.cuda.synchronize()
pool = step.bufs.pool
gates = {"shared": (pool["feat"] > 0).cpu(), "frame_disc": (pool["hid_f"] > 0).cpu(),
         "trn": [(a > 0).cpu() for a in pool["act"]], "rel_disc": [(h > 0).cpu() for h in pool["hid_r"]],
         "video_disc": (pool["hid_v"] > 0).cpu()}
plain = orc.activation_pattern(p64, xs.double(), xt.double(), beta, cfg)
lst = lambda g: [g["shared"], g["frame_disc"], *g["trn"], *g["rel_disc"], g["video_disc"]]
flips = sum((a != b).sum().item() for a, b in zip(lst(gates), lst(plain))); total = sum(t.numel() for t in lst(gates))
lp, op, gp = orc.train_step(p64, xs.double(), xt.double(), labels, beta, cfg, 0.003, train=True, gates=gates)
pinned = {"loss": rel(lossp[0], lp)}
for n, p in m.named_parameters():
    if n in gp:
        pinned["grad:" + n] = rel(p.grad, gp[n])
ref = {"loss": rel(l32, l64)}
for n, a, b in zip(names, flat(o32), flat(o64)):
    ref["out:" + n] = rel(a, b)
for n in g64:
    ref["grad:" + n] = rel(g32[n], g64[n])
print(f"# B={B}+{B} T={T} perturbed={os.environ.get('PERTURB','1')}  normwise rel err vs fp64 oracle")
print(f"# ReLU units whose on/off state differs between the tf32 CUDA forward and the fp64 oracle: {flips} of {total} ({flips/total:.2e})")
print(f"{'tensor':58s} {'cpu fp32':>10s} {'cuda fp32':>10s} {'cuda tf32':>10s} {'tf32 pinned':>12s}" + (f" {'cuda tf32x3':>12s}" if X3 else ""))
for k in res["fp32"]:
    pin = f"{pinned[k]:12.2e}" if k in pinned else f"{'':>12s}"
    x3 = f" {res['tf32x3'][k]:12.2e}" if X3 else ""
    print(f"{k:58s} {ref[k]:10.2e} {res['fp32'][k]:10.2e} {res['tf32'][k]:10.2e} {pin}{x3}")
