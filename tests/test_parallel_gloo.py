"""Host-side data-parallel logic on CPU: world_size-2 gloo (the NCCL path uses the same code)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import ta3n_oracle as orc
from ta3n_b200.parallel import GradientBucket, shard_rows


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(2)
        cfg = orc.PathConfig(num_class=5, num_segments=3, fc_dim=64, dropout_i=0.0, dropout_v=0.0)
        params = orc.init_params(cfg, seed=11)
        xs, xt, labels = orc.synthetic_batch(8, cfg, seed=5)
        sl = shard_rows(8, rank, world)
        # local shard gradients, computed by the oracle standing in for the CUDA path
        _, _, grads = orc.train_step(params, xs[sl], xt[sl], labels[sl], (0.75, 0.75, 0.5), cfg, 0.003, train=False)
        holders = []
        for name, p in params.items():
            t = torch.nn.Parameter(p.clone().float()) if p.dtype.is_floating_point else None
            if t is None:
                continue
            t.grad = grads[name].clone() if name in grads else None     # unused params keep grad None
            holders.append((name, t))
        bucket = GradientBucket([t for _, t in holders])
        bucket.allreduce_mean()
        pending = bucket.allreduce_mean(async_op=True)                   # a second (async) round: mean of equal values
        pending.wait()
        if rank == 0:
            q.put({n: (t.grad.numpy().copy() if t.grad is not None else None) for n, t in holders})
            q.put(bucket.nbytes)
    finally:
        dist.destroy_process_group()


def test_two_rank_mean_of_shard_gradients_equals_full_batch_gradient():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=240)
    nbytes = q.get(timeout=60)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    cfg = orc.PathConfig(num_class=5, num_segments=3, fc_dim=64, dropout_i=0.0, dropout_v=0.0)
    params = orc.init_params(cfg, seed=11)
    xs, xt, labels = orc.synthetic_batch(8, cfg, seed=5)
    _, _, full = orc.train_step(params, xs, xt, labels, (0.75, 0.75, 0.5), cfg, 0.003, train=False)
    used = 0
    for name, g in got.items():
        if name in full:
            used += full[name].numel()
            g = torch.from_numpy(g)
            err = (g.double() - full[name].double()).norm() / full[name].double().norm().clamp_min(1e-12)
            assert err < 1e-4 or (g - full[name]).abs().max() < 1e-7, (name, err.item())   # cancelling bias sums
        else:
            assert g is None, name                     # parameters off the path stay without grad
    assert nbytes == used * 4                          # the bucket holds exactly the used gradients


def test_shard_rows_requires_even_split():
    assert shard_rows(8, 1, 2) == slice(4, 8)
    with pytest.raises(ValueError):
        shard_rows(7, 0, 2)
