"""Where does the time of the fused step kernel go?  Runs TrainStep(mode='fused') eagerly with the per-task trace on
and prints, per task kind, count / busy time, the wall span of each stage and SM utilisation.
    python tools/step_trace.py [B] [T] [C]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ta3n_b200  # noqa: E402
from ta3n_b200.models import VideoModel  # noqa: E402
from ta3n_b200.train import TrainStep  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
T = int(sys.argv[2]) if len(sys.argv) > 2 else 5
Cn = int(sys.argv[3]) if len(sys.argv) > 3 else 12
dev = torch.device("cuda:0")
ta3n_b200.set_gemm_engine("tf32")
torch.manual_seed(1234)
m = VideoModel(Cn, "video", "trn-m", "RGB", train_segments=T, val_segments=T, add_fc=1, fc_dim=512, dropout_i=0.5,
               dropout_v=0.5, partial_bn=False, use_bn="none", ens_DA="none", use_attn="TransAttn",
               use_attn_frame="none", share_params="Y", verbose=False).to(dev).train()
g = torch.Generator().manual_seed(4321)
xs, xt = torch.randn(B, T, 2048, generator=g), torch.randn(B, T, 2048, generator=g)
labels = torch.arange(B) % Cn
step = TrainStep(m, B, B, (0.75, 0.75, 0.5), gamma=0.003, use_graph=False, mode="fused")
step.load(xs, xt, labels)
for _ in range(3):
    step.run()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
tr = step.trace(True)
flush.fill_(1)
step.run()
torch.cuda.synchronize()
t_all = tr.cpu()
n_tasks = step.step_info()[0]
t = t_all[:n_tasks]
marks = t_all[n_tasks:].double()
tag, sched, acc, done, body, synced = t[:, 0], t[:, 1], t[:, 2], t[:, 3], t[:, 4], t[:, 5]
sm = tag & 0xFFFF
typ = (tag >> 16) & 0xFF
grp = (tag >> 24) & 0xFFFFFF
mode = (tag >> 48) & 0xF
nit = (tag >> 52) & 0xFFF
t0 = int(sched[sched > 0].min())
span = (int(done.max()) - t0) / 1e3
print(f"tasks {len(t)}  span {span:.1f} us  SMs used {len(set(sm.tolist()))}")
names = {0: "gemm", 1: "row", 6: "frame", 2: "colsum_part", 3: "colsum_reduce", 4: "finish"}
busy_total = 0.0
for k, nm in names.items():
    sel = typ == k
    if sel.sum() == 0:
        continue
    dur = (done[sel] - sched[sel]).double() / 1e3
    busy_total += float(dur.sum())
    extra = ""
    if k == 0:
        ml = (acc[sel] - sched[sel]).double() / 1e3
        ep = (done[sel] - acc[sel]).double() / 1e3
        extra = f"  wait-for-accumulator avg {ml.mean():.2f} us  epilogue avg {ep.mean():.2f} us  slabs {int(nit[sel].sum())}"
    print(f"  {nm:14s} n={int(sel.sum()):5d}  busy {dur.sum():9.1f} us  avg {dur.mean():6.2f}  max {dur.max():6.2f}{extra}")
print(f"  epilogue-warp busy fraction: {busy_total / (span * len(set(sm.tolist()))):.2f}")
# per GEMM group: when did it start / end (relative), tiles, avg duration
print("  group  tiles  slabs/tile  first_start  last_end   avg_dur  acc_wait  epilogue = body + sync + publish (us)")
for gidx in sorted(set(grp[typ == 0].tolist())):
    sel = (typ == 0) & (grp == gidx)
    print(f"  {gidx:5d}  {int(sel.sum()):5d}  {float(nit[sel].double().mean()):9.1f}  {(int(sched[sel].min()) - t0) / 1e3:10.1f}  "
          f"{(int(done[sel].max()) - t0) / 1e3:9.1f}  {float((done[sel] - sched[sel]).double().mean()) / 1e3:8.2f}"
          f"  {float((acc[sel] - sched[sel]).double().mean()) / 1e3:8.2f}  {float((done[sel] - acc[sel]).double().mean()) / 1e3:8.2f}"
          f"  {float((body[sel] - acc[sel]).double().mean()) / 1e3:6.2f} {float((synced[sel] - body[sel]).double().mean()) / 1e3:6.2f}"
          f" {float((done[sel] - synced[sel]).double().mean()) / 1e3:6.2f}")
for k in (6, 2, 3):
    sel = typ == k
    if sel.sum():
        print(f"  {names[k]:14s} first_start {(int(sched[sel].min()) - t0) / 1e3:8.1f}  last_end {(int(done[sel].max()) - t0) / 1e3:8.1f}")
for kind, nm in ((0, "relpool"), (1, "heads"), (2, "relbwd")):
    sel = (typ == 1) & (mode == kind)
    if sel.sum():
        d = (done[sel] - sched[sel]).double() / 1e3
        print(f"  row:{nm:9s} n={int(sel.sum()):4d} first_start {(int(sched[sel].min()) - t0) / 1e3:8.1f}  last_end {(int(done[sel].max()) - t0) / 1e3:8.1f}"
              f"  avg {d.mean():6.2f} max {d.max():6.2f}")
# idle time: per SM, span minus busy
busy = {}
for i in range(len(t)):
    busy[int(sm[i])] = busy.get(int(sm[i]), 0.0) + (int(done[i]) - int(sched[i])) / 1e3
print(f"  per-SM busy: min {min(busy.values()):.1f} max {max(busy.values()):.1f} mean {sum(busy.values()) / len(busy):.1f} us of {span:.1f}")

# phase marks of the row tasks (warp 0 of each task): relpool [start, logits done, pooled, end], heads [start, logits, loss, end]
ok = (marks[:, 0] > 0) & (marks[:, 3] > 0)
if ok.any():
    m = marks[ok]
    print("  relpool phases (us): logits %.2f  pool %.2f  dropout+store %.2f" % (
        float((m[:, 1] - m[:, 0]).mean()) / 1e3, float((m[:, 2] - m[:, 1]).mean()) / 1e3, float((m[:, 3] - m[:, 2]).mean()) / 1e3))
    print("  heads phases (us): logits %.2f  loss %.2f  backward %.2f" % (
        float((m[:, 5] - m[:, 4]).mean()) / 1e3, float((m[:, 6] - m[:, 5]).mean()) / 1e3, float((m[:, 7] - m[:, 6]).mean()) / 1e3))
