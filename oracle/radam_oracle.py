"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): CPU restatement of the reference optimizer step.

Follows /root/reference/radam.py:44-122 (RAdam.step) and the gradient clipping the training loop applies just before
it (train.py:324-329, torch.nn.utils.clip_grad_norm_).  Pinned against the unmodified reference class by
tests/golden/radam.npz (tests/make_golden.py:radam_case).
"""
from __future__ import annotations

import math

import torch


def radam_scalars(step: int, lr: float, beta1: float, beta2: float):
    """(N_sma, step_size) of radam.py:87-107; `step` is the 1-based step counter after the increment."""
    beta2_t = beta2 ** step
    n_sma_max = 2.0 / (1.0 - beta2) - 1.0
    n_sma = n_sma_max - 2.0 * step * beta2_t / (1.0 - beta2_t)
    if n_sma >= 5:                                                      # radam.py:97 "more conservative"
        step_size = lr * math.sqrt((1 - beta2_t) * (n_sma - 4) / (n_sma_max - 4) * (n_sma - 2) / n_sma
                                   * n_sma_max / (n_sma_max - 2)) / (1 - beta1 ** step)
    else:
        step_size = lr / (1 - beta1 ** step)
    return n_sma, step_size


def radam_step(p, g, m, v, step, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
    """One RAdam update on fp32 tensors; returns (p, m, v) new tensors.  radam.py:77-122."""
    beta1, beta2 = betas
    v = v * beta2 + (1 - beta2) * g * g                                 # :78
    m = m * beta1 + (1 - beta1) * g                                     # :79
    n_sma, step_size = radam_scalars(step, lr, beta1, beta2)
    if weight_decay != 0:
        p = p + (-weight_decay * lr) * p                                # :109-112
    if n_sma >= 5:
        p = p + (-step_size) * (m / (v.sqrt() + eps))                   # :115-117
    else:
        p = p + (-step_size) * m                                        # :119
    return p, m, v


def clip_coef(grads, max_norm: float):
    """torch.nn.utils.clip_grad_norm_ (train.py:326): (total_norm, coefficient applied to every gradient)."""
    total = torch.linalg.vector_norm(torch.stack([torch.linalg.vector_norm(g.float()) for g in grads]))
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    return total, coef
