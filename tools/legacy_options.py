"""Timing of the legacy TrainStep with its stream options (cfg2, product engine).  python tools/legacy_options.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ta3n_b200  # noqa: E402
from ta3n_b200.models import VideoModel  # noqa: E402
from ta3n_b200.train import TrainStep  # noqa: E402

dev = torch.device("cuda:0")
B, T, C = 256, 5, 12
g = torch.Generator().manual_seed(4321)
xs, xt = torch.randn(B, T, 2048, generator=g), torch.randn(B, T, 2048, generator=g)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for eng in ("tf32x3", "tf32"):
    ta3n_b200.set_gemm_engine(eng)
    for kw in ({}, {"parallel_branches": True}, {"overlap_wgrad": True}, {"parallel_branches": True, "overlap_wgrad": True}):
        torch.manual_seed(1234)
        m = VideoModel(C, "video", "trn-m", "RGB", train_segments=T, val_segments=T, add_fc=1, fc_dim=512, dropout_i=0.5,
                       dropout_v=0.5, partial_bn=False, verbose=False).to(dev).train()
        step = TrainStep(m, B, B, (0.75, 0.75, 0.5), gamma=0.003, use_graph=True, mode="legacy", **kw)
        step.load(xs, xt, torch.arange(B) % C)
        for _ in range(5):
            step.run()
        torch.cuda.synchronize()
        ts = []
        for _ in range(20):
            flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            step.run()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) * 1e3)
        ts.sort()
        print(f"{eng} {kw}: median {ts[len(ts) // 2]:.1f} us  min {ts[0]:.1f} us (L2 flushed per step)", flush=True)
