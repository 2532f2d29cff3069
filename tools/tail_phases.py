"""Development: time of each phase of the per-video row task (phased executor, stand-alone tail kernel)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ta3n_b200  # noqa: E402
from ta3n_b200 import _lib  # noqa: E402
from ta3n_b200.models import VideoModel  # noqa: E402
from ta3n_b200.train import TrainStep  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = torch.device("cuda:0")
ta3n_b200.set_gemm_engine("tf32")
m = VideoModel(12, "video", "trn-m", "RGB", train_segments=5, val_segments=5, add_fc=1, fc_dim=512, dropout_i=0.5,
               dropout_v=0.5, partial_bn=False, verbose=False).to(dev).train()
g = torch.Generator().manual_seed(1)
xs, xt = torch.randn(B, 5, 2048, generator=g), torch.randn(B, 5, 2048, generator=g)
step = TrainStep(m, B, B, (0.75, 0.75, 0.5), use_graph=False, mode="phased")
step.load(xs, xt, torch.arange(B) % 12)
step.run()
buf = torch.zeros(2 * B // 4 + 1, 16, dtype=torch.int64, device=dev)
_lib.load().ta3n_debug_set_tail_trace(buf.data_ptr())
step.run()
torch.cuda.synchronize()
_lib.load().ta3n_debug_set_tail_trace(None)
t = buf.cpu()[: 2 * B // 4]
d = (t[:, 1:] - t[:, :-1]).double() / 1e3
n = int((t[0] > 0).sum())
print("phase durations (us), mean over tasks:", [round(float(d[:, i].mean()), 2) for i in range(n - 1)])
print("task total mean", float(((t[:, n - 1] - t[:, 0]).double() / 1e3).mean()))
