"""Opt-in checks of experimental kernels (not part of the default GPU suite: set TA3N_EXPERIMENTAL=1).

TA3N_FIXUP_SPLITK=1 switches the tcgen05 GEMM launches to in-kernel split-K fix-up with balanced per-group split
factors.  The flag is read once per process, so the parity tests run in a child process."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("TA3N_EXPERIMENTAL") != "1", reason="experimental kernels are opt-in")]


def test_parity_suite_with_inkernel_splitk_fixup():
    env = dict(os.environ, TA3N_FIXUP_SPLITK="1")
    env.pop("TA3N_EXPERIMENTAL", None)
    r = subprocess.run([sys.executable, "-m", "pytest", "tests/test_gpu_parity.py", "-m", "gpu", "-q", "-x", "-k",
                        "gemm_ex or golden or fused_train or full_size or mid_size"],
                       cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-2000:]
