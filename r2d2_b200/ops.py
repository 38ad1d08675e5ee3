"""Thin torch-tensor wrappers over the C-ABI kernels (device tensors in/out, current stream)."""
from __future__ import annotations

import torch

from . import _lib


def td_loss(q, qn_online, qn_target, action, n_step_reward, n_step_gamma, is_weights, learning_steps,
            want_dq: bool = True):
    """K2 (worker.py:346-359 + 268-276 + 383-390).  All inputs CUDA tensors.

    Returns (td[rows], priorities[B], loss_sum[1], rows[1] int32, dq[rows, A] | None) where
    loss = loss_sum / rows and dq = d loss_sum / d q.
    """
    rows, A = q.shape
    B = learning_steps.numel()
    dev = q.device
    assert action.dtype == torch.uint8 and learning_steps.dtype == torch.uint8
    td = torch.empty(rows, dtype=torch.float32, device=dev)
    prio = torch.empty(B, dtype=torch.float32, device=dev)
    loss_sum = torch.empty(1, dtype=torch.float32, device=dev)
    nrows = torch.empty(1, dtype=torch.int32, device=dev)
    dq = torch.empty(rows, A, dtype=torch.float32, device=dev) if want_dq else None
    p = _lib.ptr
    _lib.check(_lib.lib().r2d2_td_loss(p(q), p(qn_online), p(qn_target), p(action.reshape(-1)), p(n_step_reward),
                                       p(n_step_gamma), p(is_weights), p(learning_steps), B, A, p(td), p(prio),
                                       p(loss_sum), p(nrows), p(dq), _lib.stream_ptr()))
    return td, prio, loss_sum, nrows, dq
