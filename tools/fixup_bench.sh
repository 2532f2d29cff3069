#!/bin/bash
# A/B of the experimental knobs: in-kernel split-K fix-up, TMA L2 prefetch distance.
#   gpurun --timeout 600 -- 'bash tools/fixup_bench.sh'
set -u
for v in 0 1; do
  TA3N_FIXUP_SPLITK=$v timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('fixup=$v', round(d['ms_per_step'],4), 'ms/step', {k: d['kernel_ms_per_step'][k] for k in ('fwd_batch','shared_fc_fwd','trn_dgrad','wgrad_all','disc_dgrad','disc_fwd','relattn_fwd','relattn_dgrad','splitk_reduce') if k in d['kernel_ms_per_step']})"
done
timeout 200 python bench.py --engine tf32x3 --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('engine=tf32x3', round(d['ms_per_step'],4), 'ms/step', {k: d['kernel_ms_per_step'][k] for k in ('fwd_batch','shared_fc_fwd','trn_dgrad','wgrad_all') if k in d['kernel_ms_per_step']})"
TA3N_DESC_PREFETCH=1 timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('desc_prefetch=1', round(d['ms_per_step'],4), 'ms/step')"
for mnk in 40000000; do      # the 512x256x256 layers: video + relation discriminators, fwd and dgrad
  TA3N_SIMT_MAX_MNK=$mnk timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('simt_max_mnk=$mnk', round(d['ms_per_step'],4), 'ms/step', {k: d['kernel_ms_per_step'][k] for k in ('disc_fwd','disc_dgrad','relattn_fwd','relattn_dgrad') if k in d['kernel_ms_per_step']})"
done
for pf in 8 16 32; do
  TA3N_L2_PREFETCH=$pf timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('l2_prefetch=$pf', round(d['ms_per_step'],4), 'ms/step', {k: d['kernel_ms_per_step'][k] for k in ('fwd_batch','shared_fc_fwd','trn_dgrad','wgrad_all') if k in d['kernel_ms_per_step']})"
done
X3=1 timeout 120 python tools/parity_report.py 256 5 2>&1 | tail -50
TA3N_EXPERIMENTAL=1 timeout 900 python -m pytest tests/test_experimental.py -m gpu -q 2>&1 | tail -5
