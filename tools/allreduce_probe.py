"""Latency of the library's gradient all-reduce kernel alone (no step in front of it: no rank skew), per bucket size,
with and without the NVSwitch multicast mapping.    torchrun --nproc-per-node N tools/allreduce_probe.py"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ta3n_b200 import _lib  # noqa: E402

rank = int(os.environ["RANK"])
local = int(os.environ["LOCAL_RANK"])
world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
import torch.distributed._symmetric_memory as symm  # noqa: E402

lib = _lib.load()
n_max = 3483648
flat = symm.empty(n_max, dtype=torch.float32, device=dev)
flat.fill_(float(rank + 1))
hdl = symm.rendezvous(flat, dist.group.WORLD)
flags = symm.empty(lib.ta3n_allreduce_flag_bytes(world) // 4, dtype=torch.int32, device=dev)
flags.zero_()
fh = symm.rendezvous(flags, dist.group.WORLD)
torch.cuda.synchronize()
dist.barrier()
seq = torch.zeros(1, dtype=torch.int64, device=dev)
bufs = _lib.ptr_array([int(p) for p in hdl.buffer_ptrs])
fl = _lib.ptr_array([int(p) for p in fh.buffer_ptrs])
mc_ptr = int(getattr(hdl, "multicast_ptr", 0) or 0)
st = torch.cuda.current_stream().cuda_stream
for mc in ([mc_ptr, 0] if mc_ptr else [0]):
    for n in (n_max, n_max // 4, 4096):
        times = []
        for it in range(30):
            seq += 1
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            dist.barrier()
            torch.cuda.synchronize()
            a.record()
            _lib.check(lib.ta3n_allreduce_mean(bufs, mc or None, fl, seq.data_ptr(), rank, world, n, st))
            b.record()
            torch.cuda.synchronize()
            times.append(a.elapsed_time(b) * 1e3)
        times.sort()
        t = torch.tensor([times[len(times) // 2]], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(f"world {world}  {'multicast' if mc else 'peer ld/st'}  n = {n} floats ({n * 4 / 1e6:.2f} MB): median {t.item():.1f} us "
                  f"(max over ranks; includes the host-side skew left by dist.barrier)", flush=True)
# back-to-back inside one stream: the ranks stay in lock step, no host skew
for mc in ([mc_ptr, 0] if mc_ptr else [0]):
    n = n_max
    dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for it in range(50):
        seq += 1
        _lib.check(lib.ta3n_allreduce_mean(bufs, mc or None, fl, seq.data_ptr(), rank, world, n, st))
    b.record()
    torch.cuda.synchronize()
    if rank == 0:
        print(f"world {world}  {'multicast' if mc else 'peer ld/st'}  back to back: {a.elapsed_time(b) * 1e3 / 50:.1f} us per call "
              f"(13.9 MB, includes the seq increment kernel)", flush=True)
dist.destroy_process_group()
