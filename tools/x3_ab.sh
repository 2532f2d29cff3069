for d in 1 0; do
  echo "=== TA3N_X3_DGRAD=$d"
  TA3N_X3_DGRAD=$d X3=1 PERTURB=0 python tools/parity_report.py 256 2>&1 | grep "^grad\|^loss\|^# " | awk '{print $1, $2, $3, $4, $5, $6}'
  TA3N_X3_DGRAD=$d python - <<'PY'
import torch, sys
sys.path.insert(0, '.')
import ta3n_b200
from ta3n_b200.models import VideoModel
from ta3n_b200.train import TrainStep
from ta3n_b200 import _lib
dev = torch.device("cuda:0")
for eng in ("tf32", "tf32x3"):
    ta3n_b200.set_gemm_engine(eng)
    torch.manual_seed(1234)
    m = VideoModel(12, "video", "trn-m", "RGB", train_segments=5, val_segments=5, add_fc=1, fc_dim=512, dropout_i=0.5, dropout_v=0.5,
                   partial_bn=False, verbose=False).to(dev).train()
    g = torch.Generator().manual_seed(4321)
    xs, xt = torch.randn(256, 5, 2048, generator=g), torch.randn(256, 5, 2048, generator=g)
    step = TrainStep(m, 256, 256, (0.75, 0.75, 0.5), gamma=0.003, use_graph=True, mode="legacy")
    step.load(xs, xt, torch.arange(256) % 12)
    for _ in range(5): step.run()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(50): step.run()
    b.record(); torch.cuda.synchronize()
    print(f"legacy/{eng}: {a.elapsed_time(b) / 50 * 1e3:.1f} us/step")
    e = TrainStep(m, 256, 256, (0.75, 0.75, 0.5), gamma=0.003, use_graph=False, mode="legacy")
    e.load(xs, xt, torch.arange(256) % 12)
    for _ in range(3): e.run()
    _lib.timing_enable(True)
    for _ in range(10): e.run()
    torch.cuda.synchronize()
    rep = _lib.timing_report(); _lib.timing_enable(False)
    print("   ", {k: round(v[1] / 10 * 1e3, 1) for k, v in rep.items() if v[1] / 10 * 1e3 > 8})
PY
done
