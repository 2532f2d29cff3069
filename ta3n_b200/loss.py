"""Loss heads that follow the path in the reference's training step.

torch-op versions with the reference's semantics; they consume the outputs of VideoModel.forward
and stay differentiable through the fused operator.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def cross_entropy_soft(pred):
    """loss.py:8-12 -- mean entropy of softmax(pred)."""
    return torch.mean(torch.sum(-F.softmax(pred, dim=1) * F.log_softmax(pred, dim=1), 1))


def attentive_entropy(pred, pred_domain):
    """loss.py:15-25 -- entropy of the class prediction weighted by (1 + domain entropy)."""
    dom_ent = torch.sum(-F.softmax(pred_domain, dim=1) * F.log_softmax(pred_domain, dim=1), 1)
    cls_ent = torch.sum(-F.softmax(pred, dim=1) * F.log_softmax(pred, dim=1), 1)
    return torch.mean((1 + dom_ent) * cls_ent)


def dis_MCD(out1, out2):
    """loss.py:29-30 -- classifier discrepancy of the MCD variant (main.py:548-556): mean |softmax(out1) - softmax(out2)|."""
    return torch.mean(torch.abs(F.softmax(out1, dim=1) - F.softmax(out2, dim=1)))


def ta3n_loss(outputs, label_source, gamma=0.003, place_adv=('Y', 'Y', 'Y'), use_attn='TransAttn',
              add_loss_DA='attentive_entropy'):
    """Loss of the shipped configuration (use_target='uSv', adv_DA='RevGrad'):
    main.py:446 (source CE) + main.py:508-538 (domain CE per level) + main.py:559-562."""
    (_, out_s, _, pd_s, _, _, out_t, _, pd_t, _) = outputs
    loss = F.cross_entropy(out_s, label_source)
    per_level = []
    for lvl, flag in enumerate(place_adv):
        if flag != 'Y':
            continue
        ps, pt = pd_s[lvl].reshape(-1, 2), pd_t[lvl].reshape(-1, 2)
        dom = torch.cat([torch.zeros(ps.size(0), dtype=torch.long, device=ps.device),
                         torch.ones(pt.size(0), dtype=torch.long, device=pt.device)])
        both = torch.cat([ps, pt], 0)
        per_level.append(both)
        loss = loss + F.cross_entropy(both, dom)
    if add_loss_DA == 'attentive_entropy' and use_attn != 'none' and len(per_level) > 1:
        loss = loss + gamma * attentive_entropy(torch.cat([out_s, out_t], 0), per_level[1])
    return loss
