"""Batched GPU actor inference.

The reference steps every actor with a batch-1 CPU `Network.forward` (model.py:65-79, worker.py:526-544), which caps the
whole system at ~1,600 env-steps/s.  Here N actors share ONE forward of the learner's own kernels on a T = 1 workspace:
`r2d2_net_forward` (conv encoder -> LSTM step -> dueling head, rows b + t with b = 0, l = 1) gives Q for all actors and
`r2d2_net_state_after` returns their next (h, c).  No CPU fallback: the CUDA extension must be present."""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch

from . import _lib, config
from .learner_core import HIDDEN, DeviceLearner


class BatchedPolicy:
    def __init__(self, action_dim: int, num_actors: int, obs_shape=None, device=None):
        obs_shape = tuple(obs_shape if obs_shape is not None else config.obs_shape)
        self.N, self.A, self.C = num_actors, action_dim, obs_shape[0]
        self.core = DeviceLearner(action_dim, num_actors, 1, in_channels=self.C, max_learning=1, max_forward=0, device=device)
        d = self.core.device
        self.device = d
        N, A = num_actors, action_dim
        self._obs = torch.zeros(N, 1, self.C, 84, 84, dtype=torch.uint8, device=d)
        self._la = torch.zeros(N, 1, A, dtype=torch.uint8, device=d)
        self._lr = torch.zeros(N, 1, device=d)
        self._hidden = torch.zeros(N, 2, HIDDEN, device=d)
        self._next_hidden = torch.zeros(N, 2, HIDDEN, device=d)
        self._zeros = torch.zeros(N, dtype=torch.uint8, device=d)
        self._ones = torch.ones(N, dtype=torch.uint8, device=d)
        self._q = torch.zeros(self.core.rows_cap, A, device=d)

    # ---- weights ------------------------------------------------------------------------------------------------
    def load_state_dict(self, sd) -> None:
        self.core.online.load(sd)
        self.core.pack(0)

    def load_from_learner(self, learner_core: DeviceLearner) -> None:
        """device-to-device refresh from a learner on the same GPU (replaces Actor.update_weights, worker.py:564-566)"""
        self.core.online.flat.copy_(learner_core.online.flat)
        self.core.pack(0)

    # ---- one environment step of all actors ------------------------------------------------------------------------
    @_lib.on_device
    def step(self, obs, last_action, last_reward, hidden: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """obs u8 (N,C,84,84); last_action one-hot (N,A) (bool/u8/float) or indices (N,); last_reward (N,);
        hidden (N,2,512) = (h, c) per actor (zeros after a reset; None keeps the state of the previous step).
        Returns (q (N,A), next_hidden (N,2,512)), both on the device; next_hidden is valid until the next call."""
        d = self.device
        as_t = lambda x: x if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x))
        self._obs.copy_(as_t(obs).reshape(self.N, 1, self.C, 84, 84), non_blocking=True)
        la = as_t(last_action)
        if la.dim() == 1:                                   # action indices -> one-hot (AgentState.update, model.py:17-20)
            self._la.zero_()
            self._la.view(self.N, self.A).scatter_(1, la.to(d).long().view(-1, 1), 1)
        else:
            self._la.copy_((la.reshape(self.N, 1, self.A) != 0).to(torch.uint8), non_blocking=True)
        self._lr.copy_(as_t(last_reward).reshape(self.N, 1).float(), non_blocking=True)
        if hidden is not None:
            self._hidden.copy_(as_t(hidden).reshape(self.N, 2, HIDDEN).float(), non_blocking=True)
        else:
            self._hidden, self._next_hidden = self._next_hidden, self._hidden
        p = _lib.ptr
        core = self.core
        _lib.check(_lib.lib().r2d2_net_forward(core._h, 0, p(core.online.flat), p(self._obs), p(self._la), p(self._lr),
                                               p(self._hidden), p(self._zeros), p(self._ones), p(self._zeros), p(self._q),
                                               None, _lib.stream_ptr()))
        _lib.check(_lib.lib().r2d2_net_state_after(core._h, 0, 0, p(self._next_hidden), _lib.stream_ptr()))
        return self._q[:self.N], self._next_hidden
