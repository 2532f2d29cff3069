"""Fused training step of the path: forward + loss heads + backward as a fixed sequence of C-ABI
calls (no autograd graph, no torch ops), captured in CUDA graphs.

This is lines 418-576 of the reference's ``main.py`` for the shipped configuration
(use_target='uSv', adv_DA='RevGrad', add_loss_DA='attentive_entropy'):
    model(source, target, beta, mu, is_train=True, reverse=False)          main.py:418
    CE + 3 domain CEs + gamma * attentive_entropy                          main.py:446, 508-538, 559-562
    loss.backward()                                                        main.py:576
Gradients land directly in a flat fp32 bucket whose views are installed as ``param.grad``, so a stock
``torch.optim`` optimizer and ``clip_grad_norm_`` work unchanged (parameters the path never uses keep
``grad=None``, as in the reference).

Data parallelism (one process per GPU, videos sharded, weights replicated) replaces the reference's
``nn.DataParallel`` (main.py:79) by an all-reduce (mean) over that bucket.  Default with several ranks: the bucket
lives in symmetric memory and the library's own one-kernel all-reduce (csrc/allreduce.cuh: NVLink peer loads / NVSwitch
multicast reduction) follows the backward inside the SAME CUDA graph, then the optimizer -- one graph replay per
iteration.  ``allreduce='nccl'`` (or a system without symmetric memory) falls back to NCCL: the step is then captured
as TWO graphs split where the backward has produced the gradients of the video / relation / TRN layers (62 % of the
bytes); their all-reduce runs on NCCL's stream while the second graph (frame discriminator + shared layer backward)
computes, and only the second, smaller all-reduce is exposed.
"""
from __future__ import annotations

import ctypes as C
import os
import math
from dataclasses import dataclass
from typing import List, Optional, Sequence

import torch
import torch.distributed as dist

from . import _lib
from . import functional as TF
from ._lib import check

_P = TF._p
_N_LATE = 6     # path_parameters()[0:6] = shared layer W,b + frame discriminator W1,b1,W2,b2: produced last
_ALIGN = 64     # floats: every tensor of a flat buffer starts 256-byte aligned (vector stores, TMA operands)


@dataclass
class SGDNesterov:
    """``torch.optim.SGD(model.parameters(), lr, momentum, weight_decay, nesterov=True)`` (main.py:83) preceded
    by ``clip_grad_norm_(model.parameters(), clip_gradient)`` (main.py:578-581; None = no clipping), run as two
    kernels over the flat parameter / gradient / momentum buffers (C ABI ``ta3n_sgd_nesterov_step``)."""
    lr: float
    momentum: float = 0.9
    weight_decay: float = 1e-4
    clip_gradient: Optional[float] = 20.0


def lr_dann(lr0: float, p: float) -> float:
    """adjust_learning_rate_dann (main.py:800-802): p = training progress in [0, 1] (main.py:349)."""
    return lr0 / (1.0 + 10.0 * p) ** 0.75


def beta_dann(p: float) -> float:
    """main.py:351: the DANN schedule of the GRL coefficient; replaces every NEGATIVE entry of --beta (main.py:352)."""
    return 2.0 / (1.0 + math.exp(-10.0 * p)) - 1.0


def bucket_layout(params):
    """Order and offsets (in floats) of the path's tensors inside a flat buffer: the order in which the
    backward finishes their gradients, [video head, video disc, relation discs, TRN | frame disc, shared layer],
    every slot padded to ``_ALIGN`` floats.  Returns (order, offsets-by-index, total, early_total)."""
    order = list(range(_N_LATE, len(params))) + list(range(_N_LATE))
    offs, off, early = {}, 0, 0
    for idx in order:
        offs[idx] = off
        off += -(-params[idx].numel() // _ALIGN) * _ALIGN
        if idx == len(params) - 1:
            early = off
    return order, offs, off, early


def flatten_parameters(model) -> torch.Tensor:
    """Re-point the ``.data`` of the path's parameters at views of ONE flat fp32 buffer (bucket_layout order), so
    that the optimizer is a single pass over contiguous memory.  Values are preserved; idempotent.  The
    Parameter objects (and hence state_dict / load_state_dict / checkpoints) are unchanged."""
    params = model.path_parameters()
    order, offs, total, _ = bucket_layout(params)
    flat = getattr(model, "_ta3n_flat_params", None)
    if (flat is not None and flat.numel() == total and flat.device == params[0].device and
            all(params[i].data_ptr() == flat.data_ptr() + 4 * offs[i] for i in order)):
        return flat
    flat = torch.zeros(total, device=params[0].device, dtype=torch.float32)
    for idx in order:
        prm = params[idx]
        view = flat[offs[idx]:offs[idx] + prm.numel()].view_as(prm)
        view.copy_(prm.data)
        prm.data = view
    object.__setattr__(model, "_ta3n_flat_params", flat)
    return flat


class TrainStep:
    """Options ``overlap_wgrad`` / ``parallel_branches`` put independent parts of the backward on forked streams
    inside the captured graph.  Measured on B200: forked sub-wave tcgen05 GEMM nodes of one graph do not overlap under
    plain tf32 (tools/concurrency_probe.py: 49.9 us forked vs 50.4 us sequential for two 80-tile launches), so
    ``parallel_branches`` is off by default; ``overlap_wgrad`` defaults to on only under the tf32x3 engine, where the
    small precise weight-gradient launch hides behind the data-gradient chain (tools/legacy_options.py: -12 us)."""

    def __init__(self, model, batch_source: int, batch_target: int, beta: Sequence[float], gamma: float = 0.003,
                 place_adv: Sequence[str] = ("Y", "Y", "Y"), add_loss_DA: str = "attentive_entropy",
                 use_graph: bool = True, process_group=None, seed: int = 0x5EED, double_buffer: bool = False,
                 overlap_wgrad: Optional[bool] = None, parallel_branches: bool = False,
                 overlap_allreduce: Optional[bool] = None, graph_collectives: Optional[bool] = None,
                 optimizer: Optional[SGDNesterov] = None, mode: Optional[str] = None,
                 class_weight: Optional[torch.Tensor] = None, domain_weight: Sequence[float] = (1.0, 1.0),
                 allreduce: Optional[str] = None):
        """mode: 'legacy' (default) = the per-operator sequence (25 launches in one CUDA graph; the only mode that
        supports use_attn_frame); 'phased' = the step program as 14 launches (ta3n_step_run_phased; default when class /
        domain weights or a scheduled beta are given); 'fused' = the same program as ONE persistent kernel (C ABI
        ta3n_step_build / ta3n_step_run; plain tf32 tiles whatever engine is selected).  class_weight / domain_weight: the weights of criterion / criterion_domain
        (main.py:160-167, 204-205; fused and phased modes).  A negative beta entry selects the DANN schedule for that
        level (main.py:350-352): call set_progress(p) every step.
        allreduce (world > 1): 'peer' = this library's one-kernel all-reduce over NVLink peer / NVSwitch multicast
        memory (csrc/allreduce.cuh; the gradient bucket then lives in symmetric memory and the whole iteration --
        step, all-reduce, optimizer -- is one CUDA graph), 'nccl' = torch.distributed all_reduce between graphs;
        default: 'peer' when symmetric memory can be set up (and no NCCL overlap option was asked for), else 'nccl'."""
        if not model.training:
            raise ValueError("TrainStep needs model.train() (dropout state is fixed at construction)")
        if model.use_attn == "general" or getattr(model, "ens_DA", "none") != "none" or \
                model.frame_aggregation != "trn-m":
            # the off-path variants run through VideoModel.forward + autograd; the captured step covers
            # the shipped configurations (use_attn 'TransAttn' / 'none', one classifier)
            raise NotImplementedError("TrainStep covers frame_aggregation='trn-m', use_attn in ('TransAttn', 'none') and "
                                      "ens_DA='none'; train the other variants with model(...) + loss.backward()")
        self.model = model
        self.params = model.path_parameters()
        dev = self.params[0].device
        if dev.type != "cuda":
            raise _lib.Ta3nError("TrainStep needs the model on a CUDA device; there is no CPU path")
        self.device = dev
        self.Bs, self.Bt = int(batch_source), int(batch_target)
        self.T, self.D = model.train_segments, model.feature_dim
        self.M, self.R = self.Bs + self.Bt, self.T - 1
        self.C = model.fc_classifier_video_source.weight.shape[0]
        self.gamma = float(gamma)
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.flags = (1 if place_adv[0] == "Y" else 0) | (2 if place_adv[1] == "Y" else 0) | \
                     (4 if place_adv[2] == "Y" else 0)
        if add_loss_DA == "attentive_entropy" and model.use_attn != "none":
            if place_adv[0] != "Y" or place_adv[1] != "Y":
                # main.py:560 indexes pred_domain_all[1], the video level only when both are on
                raise NotImplementedError("attentive_entropy needs place_adv[0] == place_adv[1] == 'Y'")
            self.flags |= 8
        # split the step in two graphs around the first gradient bucket only when there is something to overlap
        if mode is None:
            # 'legacy' (the per-operator sequence) is the fastest executor measured so far and honours the selected
            # GEMM engine; the features only the step program has (class / domain weights, the DANN beta
            # schedule) select 'phased', which honours the engine too.  'fused' (one persistent kernel, plain tf32
            # tiles only) is opt-in.
            needs_step = class_weight is not None or any(float(b) < 0 for b in beta) or \
                tuple(float(w) for w in domain_weight) != (1.0, 1.0)
            mode = "phased" if (needs_step and model.use_attn_frame == "none") else "legacy"
            mode = os.environ.get("TA3N_STEP_MODE", mode)
        if mode not in ("fused", "phased", "legacy"):
            raise ValueError(f"unknown TrainStep mode {mode!r}")
        if mode != "legacy" and model.use_attn_frame != "none":
            raise NotImplementedError("the fused step does not cover use_attn_frame; use mode='legacy'")
        if overlap_wgrad is None:
            # measured at cfg2 (tools/legacy_options.py): the weight-gradient launches on a forked stream of the graph
            # gain 12 us per step under the tf32x3 engine (its small precise launch overlaps the data-gradient chain)
            # and lose 16 us under plain tf32
            overlap_wgrad = mode == "legacy" and _lib.get_gemm_engine() == "tf32x3" and not overlap_allreduce
        if mode != "legacy" and (overlap_wgrad or parallel_branches or overlap_allreduce or graph_collectives):
            raise ValueError("overlap_wgrad / parallel_branches / overlap_allreduce / graph_collectives are options "
                             "of mode='legacy' (the fused step is a single kernel)")
        self.mode = mode
        self.beta_spec = [float(b) for b in beta]
        if mode == "legacy" and any(b < 0 for b in self.beta_spec):
            raise ValueError("negative beta (= DANN schedule, main.py:350-352) needs mode='fused' or 'phased': the "
                             "legacy sequence bakes beta into the captured graph")
        if class_weight is not None and mode == "legacy":
            raise ValueError("class_weight needs mode='fused' or 'phased'")
        self._overlap_requested = bool(overlap_allreduce)
        self.split = (self.world > 1) if overlap_allreduce is None else bool(overlap_allreduce)
        if model.use_attn_frame != "none" or mode != "legacy":
            self.split = False     # frame attention couples the TRN and frame-discriminator gradients
        # graph_collectives=True captures the two NCCL all-reduces INSIDE the step's graph.  It works and is
        # marginally faster (N=2: 0.371 vs 0.378 ms/step) but process-group teardown then hangs while the graphs
        # are alive (observed on torch 2.11 / NCCL 2.28), so it is opt-in; the default is two graphs with the
        # early bucket's all-reduce issued eagerly between them (overlapping the second graph).
        self.graph_collectives = False if graph_collectives is None else bool(graph_collectives)
        self.collectives_captured = False

        # flat gradient bucket, laid out in the order the backward finishes the gradients:
        #   [ video head, video disc, relation discs, TRN | frame disc, shared layer ]
        # views are installed as .grad; the parameters live in a twin flat buffer (same offsets)
        self.flat_param = flatten_parameters(model)
        order, offs, n, self.early_numel = bucket_layout(self.params)
        self.flat_grad = self._alloc_gradient_bucket(n, allreduce)
        if self.ar is not None and overlap_allreduce is None:
            self.split = False      # the library's own all-reduce runs inside the one graph: nothing to split around
        self.grad_views: List[Optional[torch.Tensor]] = [None] * len(self.params)
        for idx in order:
            p = self.params[idx]
            self.grad_views[idx] = self.flat_grad[offs[idx]:offs[idx] + p.numel()].view_as(p)
            p.grad = self.grad_views[idx]
        self.bucket_early = self.flat_grad[:self.early_numel]
        self.bucket_late = self.flat_grad[self.early_numel:]

        # optimizer state: momentum buffers, device-resident learning rate, {norm, coef} stats
        self.opt = optimizer
        if optimizer is not None:
            self.momentum_buf = torch.zeros_like(self.flat_grad)
            self.lr_dev = torch.full((1,), float(optimizer.lr), device=dev, dtype=torch.float32)
            self._lr_host = torch.full((1,), float(optimizer.lr), dtype=torch.float32).pin_memory()
            self.grad_stats = torch.zeros(2, device=dev, dtype=torch.float32)     # [total_norm, clip_coef]
            self.opt_ws = torch.zeros(max(1, _lib.load().ta3n_sgd_workspace_bytes() // 4), device=dev,
                                      dtype=torch.float32)
            # parameters the configured losses never reach keep grad = None in the reference, and torch.optim.SGD then
            # leaves them alone (no weight decay, no momentum; main.py:83): mask their slots out of the fused update
            R = self.R
            idle = []
            if not (self.flags & 4) and model.use_attn_frame == "none":
                idle += [2, 3, 4, 5]                                  # frame discriminator
            if not (self.flags & 1) and model.use_attn == "none":
                idle += list(range(6 + 2 * R, 6 + 6 * R))             # relation discriminators
            if not (self.flags & 2) and not (self.flags & 8):
                idle += list(range(len(self.params) - 4, len(self.params)))     # video discriminator
            self.active_mask = None
            if idle:
                self.active_mask = torch.ones_like(self.flat_grad)
                for idx in idle:
                    self.active_mask[offs[idx]:offs[idx] + -(-self.params[idx].numel() // _ALIGN) * _ALIGN] = 0.0

        f32 = dict(device=dev, dtype=torch.float32)
        # input slots: one, or two for prefetching the next mini-batch while this one computes
        self.n_slots = 2 if double_buffer else 1
        # per slot: source features, target features, source labels, {real source rows, real target rows}
        self.slots = [(torch.zeros(self.Bs, self.T, self.D, **f32), torch.zeros(self.Bt, self.T, self.D, **f32),
                       torch.zeros(self.Bs, device=dev, dtype=torch.int64),
                       torch.tensor([self.Bs, self.Bt], device=dev, dtype=torch.int32)) for _ in range(self.n_slots)]
        self._valid_host = [torch.tensor([self.Bs, self.Bt], dtype=torch.int32).pin_memory()
                            for _ in range(self.n_slots)]
        self.active = 0
        self.xs, self.xt, self.labels, self.valid = self.slots[0]
        self.copy_stream = torch.cuda.Stream(device=dev) if double_buffer else None
        self.ready = [None] * self.n_slots          # event: slot filled
        self.consumed = [None] * self.n_slots       # event: last step that read the slot has finished
        self.loss = torch.zeros(1, **f32)
        self.g_video = torch.zeros(self.M, self.C, **f32)
        self.g_rel = torch.zeros(self.M, self.R, 2, **f32)
        self.g_dom = torch.zeros(self.M, 2, **f32)
        self.g_frame = torch.zeros(self.M * self.T, 2, **f32)
        self.step_counter = torch.zeros(1, device=dev, dtype=torch.int64)
        self.bufs = TF.Buffers(dev, persistent=True)
        self.loss_ws = self.bufs.workspace("loss", _lib.load().ta3n_loss_workspace_bytes(self.M))

        di, dv = float(model.dropout_rate_i), float(model.dropout_rate_v)
        # independent dropout masks per data-parallel rank, like the reference's DataParallel replicas
        rank = dist.get_rank(process_group) if dist.is_initialized() else 0
        seed = (int(seed) ^ (rank * 0x9E3779B97F4A7C15)) & (2 ** 63 - 1)
        # GRL coefficients in device memory (fused / phased): rescheduled per step without re-capturing
        self.beta_dev = torch.tensor([max(b, 0.0) for b in self.beta_spec], device=dev, dtype=torch.float32)
        self._beta_host = torch.tensor([max(b, 0.0) for b in self.beta_spec], dtype=torch.float32).pin_memory()
        self.class_weight = None if class_weight is None else \
            class_weight.detach().to(device=dev, dtype=torch.float32).contiguous()
        if self.class_weight is not None and self.class_weight.numel() != self.C:
            raise ValueError("class_weight must have one entry per class")
        self.domain_weight = (float(domain_weight[0]), float(domain_weight[1]))
        self.spec = TF.PathSpec(
            num_segments=self.T, beta=tuple(max(b, 0.0) for b in self.beta_spec), mu=0.0, reverse=False,
            use_attn=model.use_attn != "none", use_attn_frame=model.use_attn_frame != "none",
            drop_i=TF.DropSpec(p=di, seed=seed, step=self.step_counter) if di > 0 else TF.DropSpec(),
            drop_v=TF.DropSpec(p=dv, seed=seed ^ 0x9E3779B9, step=self.step_counter) if dv > 0 else TF.DropSpec())
        self.outputs = None
        self.branch_stream = torch.cuda.Stream(device=dev) if parallel_branches else None
        self.overlap_wgrad = bool(overlap_wgrad)
        self.side_stream = torch.cuda.Stream(device=dev) if self.overlap_wgrad else None
        # measured at N=2 (profiles/r2_bench_n2_early_allreduce.txt): 0.448 ms/step with the early bucket reduced on a
        # forked stream against 0.432 with one all-reduce behind the step -- the collective is latency / rank-skew bound
        # (a second kernel pays the fixed ~25 us again and competes for SMs), so the split is opt-in
        self.early_ar = os.environ.get("TA3N_EARLY_ALLREDUCE", "0") == "1"
        self.ar_stream = torch.cuda.Stream(device=dev) if (self.overlap_wgrad and self.ar is not None) else None
        self.launches_per_step = 0               # kernels of libta3n_sm100.so per step (counted at capture)
        self.use_graph = bool(use_graph)
        self.graphs = [None] * self.n_slots      # per input slot: (graph_a, graph_b or None)
        self.step_descs = [None] * self.n_slots  # fused / phased: ta3n_step_desc per input slot (+ keep-alives)
        self.step_handles = [None] * self.n_slots
        if self.mode != "legacy":
            for slot in range(self.n_slots):
                self._build_step(slot)
        if use_graph:
            for slot in range(self.n_slots):
                self.active = slot
                self.xs, self.xt, self.labels, self.valid = self.slots[slot]
                self.graphs[slot] = self._capture()
            self.active = 0
            self.xs, self.xt, self.labels, self.valid = self.slots[0]

    # -- gradient bucket / all-reduce ------------------------------------------------------------------
    def _alloc_gradient_bucket(self, n, allreduce):
        """The flat gradient bucket.  With several ranks it is allocated in SYMMETRIC memory (same size on every rank,
        peer-mapped over NVLink, multicast-mapped through the NVSwitch when available) so that the library's own
        all-reduce kernel can read and write every rank's copy (ta3n_allreduce_mean)."""
        import os
        self.ar = None
        legacy_nccl = self.mode == "legacy" and (self._overlap_requested or self.graph_collectives)
        want = allreduce or os.environ.get("TA3N_ALLREDUCE") or ("nccl" if legacy_nccl else "peer")
        if want not in ("peer", "nccl"):
            raise ValueError(f"allreduce must be 'peer' or 'nccl', got {want!r}")
        if self.world == 1 or want == "nccl":
            return torch.zeros(n, device=self.device, dtype=torch.float32)
        try:
            import torch.distributed._symmetric_memory as symm
            group = self.group if self.group is not None else dist.group.WORLD
            lib = _lib.load()
            flat = symm.empty(n, dtype=torch.float32, device=self.device)
            flat.zero_()
            hdl = symm.rendezvous(flat, group)
            self._ar_flag_bytes = lib.ta3n_allreduce_flag_bytes(self.world)
            flags = symm.empty(2 * self._ar_flag_bytes // 4, dtype=torch.int32, device=self.device)      # two call slots
            flags.zero_()
            fh = symm.rendezvous(flags, group)
            torch.cuda.synchronize()
            dist.barrier(group=self.group)              # every rank's flags are zero before anybody signals
            mc = int(getattr(hdl, "multicast_ptr", 0) or 0)
            # two ranks: plain peer loads / stores are faster than the switch reduction (50 vs 61 us back to back,
            # 13.9 MB; at eight ranks 80 vs 67: tools/allreduce_probe.py, profiles/r2_allreduce_probe_n*.txt)
            if os.environ.get("TA3N_ALLREDUCE_NO_MULTICAST") == "1" or (self.world <= 2 and
                                                                         os.environ.get("TA3N_ALLREDUCE_MULTICAST") != "1"):
                mc = 0
            self.ar = dict(bufs=[int(p) for p in hdl.buffer_ptrs], flags=[int(p) for p in fh.buffer_ptrs],
                           mc=mc, rank=int(hdl.rank), world=int(hdl.world_size), keep=(flat, flags, hdl, fh))
            return flat
        except Exception as e:      # no symmetric memory on this system / build: NCCL between graphs
            if allreduce == "peer":
                raise
            import warnings
            warnings.warn(f"ta3n_b200: symmetric-memory all-reduce unavailable ({type(e).__name__}: {e}); using NCCL")
            self.ar = None
            return torch.zeros(n, device=self.device, dtype=torch.float32)

    def _enqueue_allreduce(self, lo=0, hi=None, slot=0, stream=None):
        """Mean over the ranks of floats [lo, hi) of the gradient bucket on `stream` (default: the current one;
        graph-capturable).  `slot` selects one of the two flag regions, so that two calls of one step -- the early
        bucket on a forked stream, the late one behind the backward -- never share a barrier flag."""
        a = self.ar
        hi = self.flat_grad.numel() if hi is None else hi
        check(_lib.load().ta3n_allreduce_mean(
            _lib.ptr_array([p + 4 * lo for p in a["bufs"]]), (a["mc"] + 4 * lo) if a["mc"] else None,
            _lib.ptr_array([p + slot * self._ar_flag_bytes for p in a["flags"]]), _P(self.step_counter),
            a["rank"], a["world"], hi - lo, stream if stream is not None else TF._stream()))

    # -- the fused step (C ABI ta3n_step_*) ------------------------------------------------------------
    def _build_step(self, slot):
        """Describe the step for input slot `slot` (ta3n_step_desc) and, in fused mode, build its task graph."""
        lib = _lib.load()
        xs, xt, labels, valid = self.slots[slot]
        R, M, T = self.R, self.M, self.T
        (w_sh, b_sh), (w1f, b1f, w2f, b2f), trn_w, trn_b, r_w1, r_b1, r_w2, r_b2, (w_c, b_c), \
            (w1v, b1v, w2v, b2v) = TF._split_params(self.params, R)
        (dw_sh, db_sh), (dw1f, db1f, dw2f, db2f), dtrn_w, dtrn_b, dr_w1, dr_b1, dr_w2, dr_b2, (dw_c, db_c), \
            (dw1v, db1v, dw2v, db2v) = TF._split_params(self.grad_views, R)
        F, H = w_sh.shape[0], trn_w[0].shape[0]
        rs = TF.relation_set(T)
        new = self.bufs.get
        bufs = dict(feat=new("feat", M * T, F), hid_f=new("hid_f", M * T, F), pred_frame=new("pred_frame", M * T, 2),
                    act=new("act", rs.n_rel, M, H), feat_rel=new("feat_rel", M, R, H), hid_r=new("hid_r", R, M, H),
                    pred_rel=new("pred_rel", M, R, 2), attn=new("attn", M, R), feat_video=new("feat_video", M, H),
                    dropped=new("dropped", M, H), pred_video=new("pred_video", M, self.C), hid_v=new("hid_v", M, H),
                    pred_dom=new("pred_dom_video", M, 2))
        d = _lib.StepDesc()
        d.Bs, d.Bt, d.T, d.D, d.F, d.H, d.C = self.Bs, self.Bt, T, self.D, F, H, self.C
        d.use_attn = int(self.spec.use_attn)
        d.loss_flags = self.flags
        d.gamma = self.gamma
        d.domain_weight[0], d.domain_weight[1] = self.domain_weight
        d.class_weight = _P(self.class_weight)
        d.beta_dev = _P(self.beta_dev)
        d.tab = C.pointer(rs.ctable)
        d.x_src, d.x_tgt, d.labels, d.valid_rows = _P(xs), _P(xt), _P(labels), _P(valid)
        di, dv = self.spec.drop_i.cstruct(), self.spec.drop_v.cstruct()
        if di is not None:
            d.drop_i = di
        if dv is not None:
            d.drop_v = dv
        keep = []          # host pointer arrays must outlive every call that reads the descriptor

        def arr(ts):
            a = _lib.ptr_array([_P(t) for t in ts])
            keep.append(a)
            return a

        d.W_sh, d.b_sh, d.W1f, d.b1f, d.W2f, d.b2f = map(_P, (w_sh, b_sh, w1f, b1f, w2f, b2f))
        d.W_trn_host, d.b_trn_host = arr(trn_w), arr(trn_b)
        d.W1r_host, d.b1r_host, d.W2r_host, d.b2r_host = arr(r_w1), arr(r_b1), arr(r_w2), arr(r_b2)
        d.Wc, d.bc, d.W1v, d.b1v, d.W2v, d.b2v = map(_P, (w_c, b_c, w1v, b1v, w2v, b2v))
        d.dW_sh, d.db_sh, d.dW1f, d.db1f, d.dW2f, d.db2f = map(_P, (dw_sh, db_sh, dw1f, db1f, dw2f, db2f))
        d.dW_trn_host, d.db_trn_host = arr(dtrn_w), arr(dtrn_b)
        d.dW1r_host, d.db1r_host, d.dW2r_host, d.db2r_host = arr(dr_w1), arr(dr_b1), arr(dr_w2), arr(dr_b2)
        d.dWc, d.dbc, d.dW1v, d.db1v, d.dW2v, d.db2v = map(_P, (dw_c, db_c, dw1v, db1v, dw2v, db2v))
        for k, t in bufs.items():
            setattr(d, k, _P(t))
        d.loss = _P(self.loss)
        d.step_counter = _P(self.step_counter)
        ws_bytes = lib.ta3n_step_workspace_bytes(C.byref(d))
        if ws_bytes == 0:
            raise _lib.Ta3nError("ta3n_step_workspace_bytes: " + (lib.ta3n_last_error() or b"?").decode())
        ws = self.bufs.workspace("step", ws_bytes)          # shared by the input slots (they never run concurrently)
        d.workspace, d.workspace_bytes = _P(ws), ws.numel()
        self.step_descs[slot] = (d, keep, rs)
        self.step_bufs = bufs
        if self.mode == "fused":
            plan = self.bufs.workspace(f"step_plan{slot}", lib.ta3n_step_plan_bytes(C.byref(d)))
            handle = C.create_string_buffer(_lib.STEP_HANDLE_BYTES)
            check(lib.ta3n_step_build(C.byref(d), _P(plan), plan.numel(), handle))
            self.step_handles[slot] = handle
        self.outputs = (bufs["feat"].view(M, T, F), bufs["pred_frame"].view(M, T, 2), bufs["attn"], bufs["pred_rel"],
                        bufs["feat_video"], bufs["pred_video"], bufs["pred_dom"])

    def step_info(self):
        """(tasks, arrival counters, GEMM tiles) of the fused step's task graph."""
        n = [C.c_int(), C.c_int(), C.c_int()]
        check(_lib.load().ta3n_step_info(self.step_handles[self.active], *[C.byref(x) for x in n]))
        return tuple(x.value for x in n)

    def trace(self, enable: bool = True):
        """Fused mode, eager or before capture: record {SM, scheduled, accumulator ready, done} per task of the step
        kernel (ta3n_step_set_trace).  Returns the device tensor (n_tasks, 8) int64 the kernel fills."""
        lib = _lib.load()
        n_tasks = self.step_info()[0]
        if not hasattr(self, "_trace_buf"):
            self._trace_buf = torch.zeros(n_tasks + (self.M + 7) // 8, 8, device=self.device, dtype=torch.int64)
        for h in self.step_handles:
            if h is not None:
                check(lib.ta3n_step_set_trace(h, _P(self._trace_buf) if enable else None))
        return self._trace_buf

    def set_beta(self, beta: Sequence[float]):
        """New GRL coefficients {relation, video, frame} for the following steps (fused / phased modes): one
        12-byte async copy, no re-capture."""
        if self.mode == "legacy":
            raise ValueError("set_beta needs mode='fused' or 'phased'")
        for i in range(3):
            self._beta_host[i] = float(beta[i])
        self.beta_dev.copy_(self._beta_host, non_blocking=True)

    def set_progress(self, p: float, lr0: Optional[float] = None):
        """Per-step schedules of main.py for training progress p in [0, 1] (main.py:349): every NEGATIVE entry of
        the configured beta takes the DANN value 2/(1+exp(-10p))-1 (main.py:350-352); with lr0 the learning rate
        follows adjust_learning_rate_dann (main.py:800-802)."""
        if any(b < 0 for b in self.beta_spec):
            bd = beta_dann(p)
            self.set_beta([bd if b < 0 else b for b in self.beta_spec])
        if lr0 is not None:
            self.set_lr(lr_dann(lr0, p))

    def install_grads(self):
        """(Re-)install the flat-bucket views as ``param.grad``.  ``optimizer.zero_grad()`` (set_to_none=True by
        default) or another TrainStep / autograd backward on the same model detaches them; run() re-installs them,
        so a stock ``zero_grad(); step.run(); optimizer.step()`` loop sees the gradients this step wrote."""
        for p, view in zip(self.params, self.grad_views):
            if view is not None and p.grad is not view:
                p.grad = view

    # -- the fixed launch sequence ---------------------------------------------------------------------
    def _enqueue_optimizer(self):
        """clip_grad_norm_ + SGD-Nesterov over the flat buffers (main.py:578-583): two launches."""
        o = self.opt
        clip = float(o.clip_gradient) if o.clip_gradient is not None else 0.0
        check(_lib.load().ta3n_sgd_nesterov_step_masked(
            _P(self.flat_param), _P(self.flat_grad), _P(self.momentum_buf), self.flat_grad.numel(),
            _P(self.lr_dev), float(o.momentum), float(o.weight_decay), clip, _P(self.opt_ws),
            self.opt_ws.numel() * 4, _P(self.grad_stats), _P(self.active_mask), TF._stream()))

    def set_lr(self, lr: float):
        """Per-step learning-rate schedules (main.py:800-802): one 4-byte async copy, no re-capture."""
        if self.opt is None:
            raise ValueError("set_lr needs TrainStep(optimizer=SGDNesterov(...))")
        self._lr_host[0] = float(lr)
        self.lr_dev.copy_(self._lr_host, non_blocking=True)
        self.opt.lr = float(lr)

    def _enqueue(self, at_split=None, optimizer=False):
        """Enqueue the whole step on the current stream.  ``at_split()`` (optional) is called at the point
        where the early gradient bucket is complete (after the TRN stage's deferred weight gradients).
        ``optimizer``: append the optimizer step (single-rank sequences; with several ranks it follows the
        all-reduce instead)."""
        lib = _lib.load()
        st = TF._stream()
        # scratch for the balanced split-K of the precise forward launches (tf32x3 engine); registered only while this
        # step is being enqueued (the captured kernels keep the address, the buffer lives as long as the TrainStep)
        scratch = self.bufs.workspace("forward_scratch", 48 << 20)
        check(lib.ta3n_set_forward_scratch(_P(scratch), scratch.numel()))
        try:
            self._enqueue_body(lib, st, at_split, optimizer)
        finally:
            check(lib.ta3n_set_forward_scratch(None, 0))

    def _enqueue_body(self, lib, st, at_split, optimizer):
        if self.mode != "legacy":
            if self.mode == "fused":
                check(lib.ta3n_step_run(self.step_handles[self.active], st))
            else:
                check(lib.ta3n_step_run_phased(C.byref(self.step_descs[self.active][0]), st))
            if self.ar is not None:
                self._enqueue_allreduce()         # same stream, same graph: step -> all-reduce -> optimizer
            if optimizer and self.opt is not None:
                self._enqueue_optimizer()
            return
        check(lib.ta3n_counter_inc(_P(self.step_counter), st))          # fresh dropout masks per step
        saved, outputs, dims = TF.path_forward(self.spec, self.xs, self.xt, self.params, self.bufs, batch_gemms=True)
        self.outputs = outputs
        _, pred_frame, _, pred_rel, _, pred_video, pred_dom, _ = outputs
        check(lib.ta3n_loss_fwd_bwd(_P(pred_video), _P(self.labels), _P(pred_rel), _P(pred_dom), _P(pred_frame),
                                    self.Bs, self.Bt, self.T, self.R, self.C, self.gamma, self.flags,
                                    _P(self.valid), _P(self.loss), _P(self.g_video), _P(self.g_rel), _P(self.g_dom),
                                    _P(self.g_frame), _P(self.loss_ws), self.loss_ws.numel(), st))
        gin = {"pred_video": self.g_video, "pred_rel": self.g_rel, "pred_dom_video": self.g_dom,
               "pred_frame": self.g_frame}
        # The data-gradient chain runs first; the weight-gradient GEMMs / bias sums it leaves behind are deferred
        # and issued as grouped launches (buffers and workspaces are persistent, so they stay valid):
        #   one batch at the end, or -- split mode -- one after the TRN stage (early bucket) and one at the end.
        main = torch.cuda.current_stream()
        side = self.side_stream
        if self.overlap_wgrad:
            flush_after = {"relation": 0, "trn": 1, "shared": 2}
        elif at_split is not None:
            flush_after = {"trn": 0, "shared": 1}
        else:
            flush_after = {"shared": 0}

        early_ar = self.ar is not None and at_split is None and self.overlap_wgrad and self.early_ar

        def stage_done(name):
            if name not in flush_after:
                return
            ws = self.bufs.workspace(f"wgrad_{flush_after[name]}", lib.ta3n_wgrad_defer_workspace_bytes())
            if self.overlap_wgrad:
                ev = torch.cuda.Event()
                ev.record(main)
                side.wait_event(ev)
                check(lib.ta3n_wgrad_defer_flush(_P(ws), ws.numel(), side.cuda_stream))
                if name == "trn" and early_ar:
                    # the video / relation / TRN gradients (62 % of the bucket) are complete on the side stream:
                    # reduce them now, on a third stream, under the frame-discriminator / shared-layer backward
                    done = torch.cuda.Event()
                    done.record(side)
                    self.ar_stream.wait_event(done)
                    self._enqueue_allreduce(0, self.early_numel, slot=0, stream=self.ar_stream.cuda_stream)
            else:
                check(lib.ta3n_wgrad_defer_flush(_P(ws), ws.numel(), TF._stream()))
            if name == "trn" and at_split is not None:
                if self.overlap_wgrad:
                    main.wait_stream(side)        # the early bucket must be complete before it is all-reduced / cut
                at_split()
            if name != "shared":
                check(lib.ta3n_wgrad_defer_begin())

        check(lib.ta3n_wgrad_defer_begin())
        TF.path_backward(self.spec, dims, self.xs, self.xt, self.params, saved, gin, self.grad_views, self.bufs,
                         stage_done=stage_done, side_stream=self.branch_stream)
        if self.overlap_wgrad:
            main.wait_stream(side)            # join
        if early_ar:
            main.wait_stream(self.ar_stream)
            self._enqueue_allreduce(self.early_numel, None, slot=1)      # the late 38 %: the only exposed part
        elif self.ar is not None and at_split is None:
            self._enqueue_allreduce()         # same stream, same graph: step -> all-reduce -> optimizer
        if optimizer and self.opt is not None:
            self._enqueue_optimizer()

    def _enqueue_with_collectives(self, optimizer=False):
        """The step with both gradient all-reduces issued in place (early bucket as soon as it is complete)."""
        pending = []
        self._enqueue(at_split=lambda: pending.append(self._allreduce(self.bucket_early, async_op=True)))
        self._allreduce(self.bucket_late)
        for w in pending:
            w.wait()
        if optimizer and self.opt is not None:
            self._enqueue_optimizer()

    def _capture(self):
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            n0 = _lib.launch_count()
            if self.world > 1 and self.split:
                self._enqueue_with_collectives()       # warm-up incl. NCCL communicator set-up
            else:
                self._enqueue(at_split=(lambda: None) if self.split else None)   # warm-up: sizes every buffer
            self.launches_per_step = _lib.launch_count() - n0        # warm-up never applies the optimizer
            if self.opt is not None:
                self.launches_per_step += 2 if self.opt.clip_gradient is not None else 1
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        if self.world > 1 and self.split and self.graph_collectives:
            try:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._enqueue_with_collectives(optimizer=True)
                self.collectives_captured = True
                return (g, None)
            except Exception as e:      # NCCL capture not possible here: eager collectives between two graphs
                import warnings
                warnings.warn(f"ta3n_b200: NCCL capture failed ({type(e).__name__}: {e}); using split graphs")
                self.collectives_captured = False
                torch.cuda.synchronize()
        # single rank, or the library's own all-reduce inside the graph: the optimizer is part of the (last) graph
        inline = self.world == 1 or self.ar is not None
        if not self.split:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._enqueue(optimizer=inline)
            return (g, None)
        # two graphs: [forward .. TRN stage + early weight gradients] | [frame discriminator + shared layer]
        ga, gb = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        cap = torch.cuda.Stream(device=self.device)
        cap.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(cap):
            ga.capture_begin()

            def cut():
                ga.capture_end()
                gb.capture_begin(pool=ga.pool())

            self._enqueue(at_split=cut, optimizer=inline)
            gb.capture_end()
        torch.cuda.current_stream().wait_stream(cap)
        torch.cuda.synchronize()
        return (ga, gb)

    # -- public API ------------------------------------------------------------------------------------
    def _fill(self, slot, source, target, labels):
        """Copy a paired mini-batch into input slot `slot` on the current stream.  Fewer than (Bs, Bt) rows --
        the last batch of an epoch -- are padded the way main.py:354-372 does and masked out of every loss term
        the way main.py:421-422 does (the rows are simply left as they were: rows are independent and the
        padded ones receive zero gradient)."""
        xs, xt, lab, valid = self.slots[slot]
        ns, nt = int(source.shape[0]), int(target.shape[0])
        if not (1 <= ns <= self.Bs and 0 <= nt <= self.Bt) or labels.shape[0] != ns:
            raise ValueError(f"batch of {ns}+{nt} videos / {labels.shape[0]} labels does not fit TrainStep({self.Bs}, {self.Bt})")
        xs[:ns].copy_(source.reshape((ns,) + tuple(xs.shape[1:])), non_blocking=True)
        if nt:
            xt[:nt].copy_(target.reshape((nt,) + tuple(xt.shape[1:])), non_blocking=True)
        lab[:ns].copy_(labels, non_blocking=True)
        host = self._valid_host[slot]
        if (int(host[0]), int(host[1])) != (ns, nt):
            # the pinned pair must not change while an earlier async copy of it may be pending; the batch size
            # changes once per epoch, so a stream synchronisation here costs nothing measurable
            torch.cuda.current_stream().synchronize()
            host[0], host[1] = ns, nt
            valid.copy_(host, non_blocking=True)

    def load(self, source, target, labels):
        """Copy one paired mini-batch (host or device tensors) into the ACTIVE input slot (compute stream)."""
        self._fill(self.active, source, target, labels)

    def prefetch(self, source, target, labels):
        """double_buffer=True: copy the NEXT mini-batch into the inactive slot on the copy stream, overlapping
        the step that is running; call ``swap()`` before the ``run()`` that should consume it."""
        if self.n_slots < 2:
            raise ValueError("prefetch needs TrainStep(double_buffer=True)")
        nxt = 1 - self.active
        with torch.cuda.stream(self.copy_stream):
            if self.consumed[nxt] is not None:
                self.copy_stream.wait_event(self.consumed[nxt])      # do not overwrite inputs still being read
            self._fill(nxt, source, target, labels)
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
        self.ready[nxt] = ev

    def swap(self):
        """Make the prefetched slot the active one (the compute stream waits for its copies)."""
        self.active = 1 - self.active
        self.xs, self.xt, self.labels, self.valid = self.slots[self.active]
        if self.ready[self.active] is not None:
            torch.cuda.current_stream().wait_event(self.ready[self.active])
            self.ready[self.active] = None

    def _allreduce(self, t, async_op=False):
        # AVG = sum * 1/world inside NCCL: mean of the equal-sized shards' gradients = global-batch gradient
        return dist.all_reduce(t, op=dist.ReduceOp.AVG, group=self.group, async_op=async_op)

    def run(self):
        """forward + loss + backward (+ gradient all-reduce) (+ optimizer step when configured); returns the
        device loss tensor (1,)."""
        pending = None
        opt_done = self.opt is None
        if self.use_graph:
            ga, gb = self.graphs[self.active]
            ga.replay()
            if gb is not None:
                if self.world > 1:
                    pending = self._allreduce(self.bucket_early, async_op=True)   # overlaps graph b
                gb.replay()
            opt_done = opt_done or self.world == 1 or self.collectives_captured or self.ar is not None
        else:
            n0 = _lib.launch_count()
            self._enqueue(optimizer=self.world == 1 or self.ar is not None)
            self.launches_per_step = _lib.launch_count() - n0
            opt_done = opt_done or self.world == 1 or self.ar is not None
        if self.n_slots > 1:
            ev = torch.cuda.Event()
            ev.record()
            self.consumed[self.active] = ev
        if self.world > 1 and self.ar is None and not (self.use_graph and self.collectives_captured):
            if pending is not None:
                self._allreduce(self.bucket_late)
                pending.wait()
            else:
                self._allreduce(self.flat_grad)
        if not opt_done:
            self._enqueue_optimizer()             # after the all-reduce: every rank applies the same update
        self.install_grads()
        return self.loss

    def __call__(self, source, target, labels):
        self.load(source, target, labels)
        return self.run()
