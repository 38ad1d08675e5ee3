"""``Network`` / ``AgentState`` -- the reference's ``model`` module surface (model.py:1-150 upstream).

``Network`` keeps the reference's parameter names, shapes and state_dict order (so checkpoints,
``share_memory()`` and ``Actor.update_weights`` interoperate), and its three entry points:

* ``forward(AgentState)``      single-step actor inference on the CPU, like upstream (model.py:65-79;
                               actors are out of the learner hot path);
* ``calculate_q_`` / ``calculate_q``  the learner-side sequence passes (model.py:81-150): these run
                               on the CUDA kernels (K1) -- the module's parameters must live on a
                               CUDA device; there is no CPU fallback.
"""
from dataclasses import dataclass, field
from typing import Optional, Tuple

import torch
import torch.nn as nn

from . import config


@dataclass
class AgentState:                                       # model.py:9-24
    obs: torch.Tensor
    action_dim: int
    last_action: torch.Tensor = field(init=False)
    last_reward: torch.Tensor = torch.zeros((1, 1), dtype=torch.float32)
    hidden_state: Optional[Tuple[torch.Tensor, torch.Tensor]] = None

    def __post_init__(self):
        self.last_action = torch.zeros((1, self.action_dim), dtype=torch.float32)

    def update(self, obs, last_action, last_reward, hidden):
        self.obs = torch.from_numpy(obs).unsqueeze(0)
        onehot = torch.zeros((1, self.action_dim), dtype=torch.float32)
        onehot[0, last_action] = 1.0
        self.last_action = onehot
        self.last_reward = torch.tensor([[last_reward]], dtype=torch.float32)
        self.hidden_state = hidden


class Network(nn.Module):
    def __init__(self, action_dim, obs_shape=config.obs_shape, hidden_dim=config.hidden_dim):
        super().__init__()
        self.action_dim = action_dim
        self.obs_shape = obs_shape
        self.hidden_dim = hidden_dim
        self.max_forward_steps = config.forward_steps
        if hidden_dim != 512:
            raise ValueError("the sm_100a kernels are specialised for hidden_dim = 512 (config.hidden_dim)")
        # same module tree as model.py:39-63 => identical state_dict keys / shapes
        self.feature = nn.Sequential(
            nn.Conv2d(obs_shape[0], 32, 8, 4), nn.ReLU(True),
            nn.Conv2d(32, 64, 4, 2), nn.ReLU(True),
            nn.Conv2d(64, 64, 3, 1), nn.ReLU(True),
            nn.Flatten(),
            nn.Linear(3136, 512), nn.ReLU(True),
        )
        self.recurrent = nn.LSTM(512 + self.action_dim + 1, self.hidden_dim, batch_first=True)
        self.advantage = nn.Sequential(nn.Linear(self.hidden_dim, self.hidden_dim), nn.ReLU(True),
                                       nn.Linear(self.hidden_dim, self.action_dim))
        self.value = nn.Sequential(nn.Linear(self.hidden_dim, self.hidden_dim), nn.ReLU(True), nn.Linear(self.hidden_dim, 1))
        self._core = None
        self._core_key = None

    # ---------------------------------------------------------------- actor path (CPU, batch 1), model.py:65-79
    def forward(self, state: AgentState):
        latent = self.feature(state.obs / 255)
        recurrent_input = torch.cat((latent, state.last_action, state.last_reward), dim=1)
        _, recurrent_output = self.recurrent(recurrent_input, state.hidden_state)
        hidden = recurrent_output[0]
        adv = self.advantage(hidden)
        val = self.value(hidden)
        return val + adv - adv.mean(1, keepdim=True), recurrent_output

    # ---------------------------------------------------------------- learner path (CUDA kernels)
    def _device_core(self, batch_size, seq_frames, max_learning):
        p = next(self.parameters())
        if not p.is_cuda:
            raise RuntimeError("Network.calculate_q*/calculate_q_ run on the sm_100a kernels: move the module to a CUDA "
                               "device (there is no CPU fallback on the learner path)")
        from .learner_core import DeviceLearner
        key = (batch_size, seq_frames, max_learning, p.device)
        if self._core_key != key:
            self._core = DeviceLearner(self.action_dim, batch_size, seq_frames, in_channels=self.obs_shape[0],
                                       max_learning=max_learning, max_forward=self.max_forward_steps, device=p.device)
            self._core_key = key
        self._core.online.load(self.state_dict())
        self._core.pack(0)
        return self._core

    def _unroll(self, obs, last_action, last_reward, hidden_state, burn_in_steps, learning_steps, forward_steps, shifted):
        B, T = obs.shape[:2]
        frames = obs if obs.dtype == torch.uint8 else (obs * 255).round().to(torch.uint8)   # callers pass obs/255 (worker.py:342)
        h0, c0 = hidden_state
        hidden = torch.stack((h0.reshape(B, -1), c0.reshape(B, -1)), dim=1)                   # (B,2,H)
        lmax = int(learning_steps.max().item())
        need = int((burn_in_steps.int() + learning_steps.int() + forward_steps.int()).max().item())
        core = self._device_core(B, max(T, need), lmax)
        b = core.prepare(dict(obs=frames, last_action=last_action, last_reward=last_reward.reshape(B, T), hidden=hidden,
                              burn_in=burn_in_steps, learning=learning_steps, forward=forward_steps))
        q_learn = torch.empty(core.rows_cap, self.action_dim, device=core.device)
        q_shift = torch.empty(core.rows_cap, self.action_dim, device=core.device)
        core.forward(0, b, q_learn, q_shift)
        rows = int(learning_steps.int().sum().item())
        return (q_shift if shifted else q_learn)[:rows]

    def calculate_q_(self, obs, last_action, last_reward, hidden_state, burn_in_steps, learning_steps, forward_steps):
        """model.py:81-119: Q at the n-step-shifted positions, rows sequence-major."""
        return self._unroll(obs, last_action, last_reward, hidden_state, burn_in_steps, learning_steps, forward_steps, True)

    def calculate_q(self, obs, last_action, last_reward, hidden_state, burn_in_steps, learning_steps):
        """model.py:122-150: Q at the learning positions (forward pass only; gradients are produced by the
        learner's fused backward, r2d2_net_backward, not by autograd)."""
        zero_f = torch.zeros_like(burn_in_steps)
        return self._unroll(obs, last_action, last_reward, hidden_state, burn_in_steps, learning_steps, zero_f, False)
