"""Data-parallel TrainStep on 2 GPUs (NCCL): the averaged shard gradients equal the single-GPU gradients of
the global batch, and with an optimizer configured every rank applies the update of the global batch (the
reference's nn.DataParallel semantics, main.py:79, 576-583).  Needs >= 2 CUDA devices; skipped otherwise."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import ta3n_oracle as orc

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")]

B_LOCAL = 32


def _cfg():
    return orc.PathConfig(num_class=12, num_segments=5, fc_dim=512, dropout_i=0.0, dropout_v=0.0)


def _build(dev):
    from ta3n_b200.models import VideoModel
    cfg = _cfg()
    m = VideoModel(cfg.num_class, "video", "trn-m", "RGB", train_segments=5, val_segments=5, fc_dim=512,
                   dropout_i=0.0, dropout_v=0.0, partial_bn=False, verbose=False)
    m.load_state_dict(orc.init_params(cfg, seed=1234))
    return m.to(dev).train()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


VARIANTS = {"nccl-legacy": dict(mode="legacy", allreduce="nccl"),      # two graphs, NCCL overlapping the second
            "peer-legacy": dict(mode="legacy", allreduce="peer"),      # one graph: step + library all-reduce + optimizer
            "peer-legacy-early": dict(mode="legacy", allreduce="peer", overlap_wgrad=True),   # + early bucket reduced
            "peer-phased": dict(mode="phased", allreduce="peer")}                             #   on a forked stream


def _worker(rank, world, port, q, variant):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        import ta3n_b200
        from ta3n_b200.parallel import shard_rows
        from ta3n_b200.train import SGDNesterov, TrainStep
        ta3n_b200.set_gemm_engine("fp32")          # exact engine: the comparison below is about the collective
        xs, xt, labels = orc.synthetic_batch(world * B_LOCAL, _cfg())
        sl = shard_rows(world * B_LOCAL, rank, world)
        model = _build(dev)
        kw = VARIANTS[variant]
        step = TrainStep(model, B_LOCAL, B_LOCAL, (0.75, 0.75, 0.5), gamma=0.0, use_graph=True, **kw)
        assert (step.graphs[0][1] is not None) == (variant == "nccl-legacy")      # split graphs only around NCCL
        assert (step.ar is not None) == (kw["allreduce"] == "peer")
        for _ in range(2):
            step(xs[sl], xt[sl], labels[sl])
        torch.cuda.synchronize()
        if rank == 0:
            q.put(step.flat_grad.cpu().numpy())
        # the same with the fused optimizer: 3 iterations, then all ranks must hold identical parameters
        model2 = _build(dev)
        opt_step = TrainStep(model2, B_LOCAL, B_LOCAL, (0.75, 0.75, 0.5), gamma=0.0, use_graph=True,
                             optimizer=SGDNesterov(lr=0.05, clip_gradient=0.02), **kw)
        for _ in range(3):
            opt_step(xs[sl], xt[sl], labels[sl])
        torch.cuda.synchronize()
        mine = opt_step.flat_param.clone()
        other = mine.clone()
        dist.broadcast(other, src=0)
        assert torch.equal(mine, other), "ranks diverged after the optimizer step"
        if rank == 0:
            q.put(opt_step.flat_param.cpu().numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("variant", list(VARIANTS))
def test_two_gpu_trainstep_matches_global_batch(variant):
    import ta3n_b200
    from ta3n_b200.train import SGDNesterov, TrainStep, flatten_parameters
    ta3n_b200.set_gemm_engine("fp32")
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, variant)) for r in range(world)]
    for p in procs:
        p.start()
    got = torch.from_numpy(q.get(timeout=300))
    got_param = torch.from_numpy(q.get(timeout=300))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    dev = torch.device("cuda", 0)
    xs, xt, labels = orc.synthetic_batch(world * B_LOCAL, _cfg())
    model = _build(dev)
    # gamma=0: every loss term is a plain mean over rows, so mean-of-shard-means == global mean exactly
    ref = TrainStep(model, world * B_LOCAL, world * B_LOCAL, (0.75, 0.75, 0.5), gamma=0.0, use_graph=True,
                    mode="legacy")                   # same bucket layout as the 2-rank run
    ref(xs, xt, labels)
    torch.cuda.synchronize()
    want = ref.flat_grad.cpu()
    err = ((got.double() - want.double()).norm() / want.double().norm()).item()
    assert err < 1e-4, err          # fp32 summation order only
    # optimizer: 3 iterations on the global batch on one GPU == 3 data-parallel iterations
    model2 = _build(dev)
    p0 = flatten_parameters(model2).clone()                    # initial values in bucket order
    ref2 = TrainStep(model2, world * B_LOCAL, world * B_LOCAL, (0.75, 0.75, 0.5), gamma=0.0, use_graph=True,
                     mode="legacy", optimizer=SGDNesterov(lr=0.05, clip_gradient=0.02))
    for _ in range(3):
        ref2(xs, xt, labels)
    torch.cuda.synchronize()
    assert float(ref2.grad_stats[1]) < 1.0                      # clipping was active
    want_delta = (ref2.flat_param - p0).double().cpu()
    got_delta = got_param.double() - p0.double().cpu()
    err = ((got_delta - want_delta).norm() / want_delta.norm()).item()
    assert err < 1e-3, err


def test_nn_dataparallel_replicas_match_single_gpu():
    """The reference's own multi-GPU mode (main.py:79: nn.DataParallel, one host thread per replica): forward outputs and
    the gradients reduced onto GPU 0 equal the single-GPU ones.  Exercises the per-device kernel configuration and the
    per-thread tensor-map caches of the library with two devices driven from one process."""
    import ta3n_b200
    from ta3n_b200.loss import ta3n_loss
    ta3n_b200.set_gemm_engine("tf32x3")
    cfg = _cfg()
    xs, xt, labels = orc.synthetic_batch(2 * B_LOCAL, cfg)
    dev0 = torch.device("cuda", 0)

    def run(parallel):
        model = _build(dev0)
        net = torch.nn.DataParallel(model, [0, 1]) if parallel else model
        outs = net(xs.to(dev0), xt.to(dev0), [0.75, 0.75, 0.5], 0, is_train=True, reverse=False)
        loss = ta3n_loss(outs, labels.to(dev0), 0.003)
        loss.backward()
        torch.cuda.synchronize()
        return loss.detach().cpu(), outs[1].detach().cpu(), {k: p.grad.detach().cpu() for k, p in model.named_parameters()
                                                             if p.grad is not None}

    l1, o1, g1 = run(False)
    l2, o2, g2 = run(True)
    assert abs(l1.item() - l2.item()) <= 1e-5 * abs(l1.item())
    assert ((o1 - o2).norm() / o1.norm()).item() < 1e-4
    for k in g1:
        err = ((g1[k].double() - g2[k].double()).norm() / g1[k].double().norm().clamp_min(1e-30)).item()
        assert err < 2e-3 or (g1[k] - g2[k]).abs().max().item() < 1e-7, (k, err)
