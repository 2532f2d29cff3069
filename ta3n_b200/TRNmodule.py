"""Temporal Relation Network module -- same public classes as the reference's TRNmodule.py.

``RelationModuleMultiScale`` keeps the reference's constructor, attribute names
(``scales``, ``relations_scales``, ``subsample_scales``, ``fc_fusion_scales``) and state_dict
keys (``fc_fusion_scales.{i}.1.{weight,bias}``; TRNmodule.py:30-56) but its forward is one call
into the CUDA library (ta3n_trn_fwd / ta3n_trn_bwd) instead of ~10 gathers + GEMMs.
"""
from __future__ import annotations

import torch
from torch import nn

from . import functional as TF


class RelationModuleMultiScale(nn.Module):
    """Multi-scale temporal relations, summed per scale (TRNmodule.py:27-86).

    forward: (N, num_frames, img_feature_dim) -> (N, num_frames-1, num_bottleneck)
    """

    def __init__(self, img_feature_dim, num_bottleneck, num_frames, nonneg_input=False):
        super().__init__()
        rs = TF.relation_set(num_frames)
        self.subsample_num = 3                                   # TRNmodule.py:32
        self.img_feature_dim = img_feature_dim
        self.num_frames = num_frames
        self.scales = list(rs.scales)                            # TRNmodule.py:34
        # full combination lists, as the reference exposes them (TRNmodule.py:36-41)
        self.relations_scales = [self.return_relationset(num_frames, s) for s in self.scales]
        self.subsample_scales = [min(self.subsample_num, len(r)) for r in self.relations_scales]
        # parameter holders with the reference's key names: Sequential(ReLU, Linear, ReLU)
        self.fc_fusion_scales = nn.ModuleList(
            nn.Sequential(nn.ReLU(), nn.Linear(s * img_feature_dim, num_bottleneck), nn.ReLU())
            for s in self.scales)
        # Inside VideoModel the input is post-ReLU/dropout (>= 0), so the leading ReLU of every
        # fc_fusion (TRNmodule.py:49) is an identity in value and gradient and can be skipped.
        self.nonneg_input = bool(nonneg_input)

    def relation_weights(self):
        return [seq[1].weight for seq in self.fc_fusion_scales], [seq[1].bias for seq in self.fc_fusion_scales]

    def forward(self, input):
        ws, bs = self.relation_weights()
        return TF.trn_multiscale(input, ws, bs, relu_input=not self.nonneg_input)

    def return_relationset(self, num_frames, num_frames_relation):
        import itertools
        return list(itertools.combinations(range(num_frames), num_frames_relation))


class RelationModule(nn.Module):
    """Single-scale TRN (TRNmodule.py:6-25).  Out of scope: the reference's 'trn' aggregation
    path crashes before producing output; kept as a name for import parity."""

    def __init__(self, img_feature_dim, num_bottleneck, num_frames):
        super().__init__()
        raise NotImplementedError("frame_aggregation='trn' is not a working path of the reference; use 'trn-m'")
