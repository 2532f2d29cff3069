"""One eager training step of the product path between cudaProfilerStart/Stop, for ncu:

    ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/r2_step_full \
        python tools/profile_step.py [engine] [mode]

Every launch of the step is captured in order; the library's launch labels (call sites) are printed in the same order to
stdout, for tools/ncu_summary.py --sites."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ta3n_b200  # noqa: E402
from ta3n_b200 import _lib  # noqa: E402
from ta3n_b200.models import VideoModel  # noqa: E402
from ta3n_b200.train import TrainStep  # noqa: E402

engine = sys.argv[1] if len(sys.argv) > 1 else "tf32x3"
mode = sys.argv[2] if len(sys.argv) > 2 else "legacy"
B, T, C = 256, 5, 12
dev = torch.device("cuda:0")
ta3n_b200.set_gemm_engine(engine)
torch.manual_seed(1234)
m = VideoModel(C, "video", "trn-m", "RGB", train_segments=T, val_segments=T, add_fc=1, fc_dim=512, dropout_i=0.5,
               dropout_v=0.5, partial_bn=False, verbose=False).to(dev).train()
g = torch.Generator().manual_seed(4321)
xs, xt = torch.randn(B, T, 2048, generator=g), torch.randn(B, T, 2048, generator=g)
step = TrainStep(m, B, B, (0.75, 0.75, 0.5), gamma=0.003, use_graph=False, mode=mode)
step.load(xs, xt, torch.arange(B) % C)
for _ in range(3):
    step.run()
torch.cuda.synchronize()
_lib.timing_enable(True)           # records the call-site label of every launch (and CUDA events around it)
torch.cuda.cudart().cudaProfilerStart()
step.run()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
rep = _lib.timing_report()
_lib.timing_enable(False)
print("launches_per_step", step.launches_per_step)
print("sites", ",".join(rep.keys()))
