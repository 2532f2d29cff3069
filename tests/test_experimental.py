"""Opt-in checks of experimental kernels (not part of the default GPU suite: set TA3N_EXPERIMENTAL=1).

Knobs: TA3N_FIXUP_SPLITK=1 in-kernel split-K fix-up with balanced per-group split
factors; TA3N_L2_PREFETCH=<slabs> TMA L2 prefetch distance; TA3N_DESC_PREFETCH=1 tensor-map descriptor prefetch;
TA3N_SIMT_MAX_MNK=<M*N*K> small GEMMs on the fp32 SIMT engine; TA3N_TEST_TF32X3=1 adds the experimental 'tf32x3'
engine (three tf32 MMAs per step on hi/lo operand splits) to the engines the parity tests run, at the fp32 tolerances."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("TA3N_EXPERIMENTAL") != "1", reason="experimental kernels are opt-in")]


KNOBS = [{"TA3N_FIXUP_SPLITK": "1"}, {"TA3N_L2_PREFETCH": "16"}, {"TA3N_DESC_PREFETCH": "1"},
         {"TA3N_SIMT_MAX_MNK": "40000000"}, {"TA3N_TEST_TF32X3": "1"},
         {"TA3N_FIXUP_SPLITK": "1", "TA3N_L2_PREFETCH": "16", "TA3N_DESC_PREFETCH": "1", "TA3N_SIMT_MAX_MNK": "40000000"}]


@pytest.mark.parametrize("knobs", KNOBS, ids=lambda k: "+".join(sorted(k)))
def test_parity_suite_under_experimental_knobs(knobs):
    """The knobs are read once per process, so the parity tests run in a child process."""
    env = dict(os.environ, **knobs)
    env.pop("TA3N_EXPERIMENTAL", None)
    r = subprocess.run([sys.executable, "-m", "pytest", "tests/test_gpu_parity.py", "-m", "gpu", "-q", "-x", "-k",
                        "gemm_ex or golden or fused_train or full_size or mid_size"],
                       cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-2000:]
