#!/bin/bash
# compute-sanitizer over a small forward+backward of the path (memcheck, racecheck, synccheck, initcheck).
# Run on a GPU box:  bash tools/sanitize.sh > gpurun_out/sanitizer.txt 2>&1
set -u
cd "$(dirname "$0")/.."
for tool in memcheck racecheck synccheck; do
  for eng in fp32 tf32; do
    echo "=== compute-sanitizer --tool $tool  (engine $eng) ==="
    TA3N_ENGINE=$eng timeout 600 compute-sanitizer --tool $tool --error-exitcode 7 \
      python -c "
import os, torch, ta3n_b200
from oracle import ta3n_oracle as orc
from ta3n_b200.models import VideoModel
from ta3n_b200.train import TrainStep
ta3n_b200.set_gemm_engine(os.environ['TA3N_ENGINE'])
cfg = orc.PathConfig(num_class=12, num_segments=5, fc_dim=512)
torch.manual_seed(0)
m = VideoModel(12, 'video', 'trn-m', 'RGB', train_segments=5, val_segments=5, fc_dim=512, partial_bn=False, verbose=False).cuda().train()
xs, xt, y = orc.synthetic_batch(24, cfg)
step = TrainStep(m, 24, 16, (0.75, 0.75, 0.5), use_graph=False)
print('loss', step(xs, xt[:16], y).item())
" 2>&1 | grep -E "ERROR SUMMARY|loss|Error|error|Race|hazard" | head -12
    echo "exit=$?"
  done
done
