"""Host-side checks of the fused step's task graph (no GPU): ta3n_step_describe builds the graph for a descriptor
with placeholder pointers and simulates a scheduler -- every task must become runnable (acyclic dependencies,
every awaited arrival count reached)."""
import ctypes as C
import re

import pytest

from ta3n_b200 import _lib
from ta3n_b200 import functional as TF


def _desc(Bs, Bt, T, C_, F=512, H=256, D=2048, drop=0.5):
    d = _lib.StepDesc()
    d.Bs, d.Bt, d.T, d.D, d.F, d.H, d.C = Bs, Bt, T, D, F, H, C_
    d.use_attn, d.loss_flags, d.gamma = 1, 15, 0.003
    d.domain_weight[0], d.domain_weight[1] = 1.0, 1.0
    rs = TF.relation_set(T)
    d.tab = C.pointer(rs.ctable)
    keep = [rs]
    fake = 1 << 30                      # never dereferenced by the host-only planner; 256-byte aligned

    def nxt():
        nonlocal fake
        fake += 1 << 26
        return fake

    R = T - 1
    for name, ctype in d._fields_:
        if ctype is C.c_void_p and name not in ("class_weight", "valid_rows"):
            setattr(d, name, nxt())
        elif ctype == C.POINTER(C.c_void_p):
            arr = (C.c_void_p * R)(*[nxt() for _ in range(R)])
            keep.append(arr)
            setattr(d, name, arr)
    d.drop_i.p = drop
    d.drop_i.seed = 1
    d.drop_i.keep = None
    d.drop_i.step_dev = nxt()
    d.drop_v.p = drop
    d.drop_v.seed = 2
    d.drop_v.keep = None
    d.drop_v.step_dev = nxt()
    d.workspace_bytes = 1 << 40
    return d, keep


@pytest.mark.parametrize("Bs,Bt,T,C_", [(256, 256, 5, 12), (512, 512, 5, 30), (128, 128, 9, 12), (8, 8, 5, 5),
                                        (3, 1, 5, 7), (60, 51, 3, 11), (130, 127, 5, 12)])
def test_task_graph_is_schedulable(Bs, Bt, T, C_):
    lib = _lib.load()
    d, keep = _desc(Bs, Bt, T, C_)
    buf = C.create_string_buffer(1024)
    n = lib.ta3n_step_describe(C.byref(d), buf, 1024)
    assert n > 0, lib.ta3n_last_error()
    text = buf.value.decode()
    fields = dict(re.findall(r"(\w+) (\d+)", text))
    assert int(fields["unsatisfiable_waits"]) == 0, text
    M = Bs + Bt
    assert int(fields["row"]) == 3 * sum(-(-min(128, M - b) // 8) for b in range(0, M, 128)), text
    assert int(fields["frame"]) == sum(-(-min(128, M * T - b) // 32) for b in range(0, M * T, 128)), text
    assert int(fields["gemm_tiles"]) > 0 and int(fields["tasks"]) < 20000
