"""Run a few eager fused steps (for ncu / compute-sanitizer captures).  python tools/step_run.py [B] [T] [C] [n]"""
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
n = int(sys.argv[4]) if len(sys.argv) > 4 else 5
mode = sys.argv[5] if len(sys.argv) > 5 else "fused"
dev = torch.device("cuda:0")
ta3n_b200.set_gemm_engine("tf32")
torch.manual_seed(1234)
m = VideoModel(Cn, "video", "trn-m", "RGB", train_segments=T, val_segments=T, add_fc=1, fc_dim=512, dropout_i=0.5,
               dropout_v=0.5, partial_bn=False, use_bn="none", ens_DA="none", use_attn="TransAttn",
               use_attn_frame="none", share_params="Y", verbose=False).to(dev).train()
g = torch.Generator().manual_seed(4321)
xs, xt = torch.randn(B, T, 2048, generator=g), torch.randn(B, T, 2048, generator=g)
labels = torch.arange(B) % Cn
step = TrainStep(m, B, B, (0.75, 0.75, 0.5), gamma=0.003, use_graph=False, mode=mode)
step.load(xs, xt, labels)
for _ in range(n):
    step.run()
torch.cuda.synchronize()
print("loss", float(step.loss))
