"""PyTorch-CPU fp32 restatement of the reference learner update (TEST INFRASTRUCTURE).

Restates, in plain functional torch on the CPU, what one iteration of the
reference's ``Learner.run`` computes (``worker.py:330-369``) together with the
network passes it calls (``model.py:81-150``).  It exists to (1) check the CUDA
path stage by stage (it exposes every intermediate: latent, hidden states,
Q-values, targets, gradients), and (2) be timed as the CPU baseline.

Parameter names are the reference ``Network.state_dict()`` keys
(``model.py:39-63``) so weights can be moved between the two verbatim.

Structure mirrors the reference on purpose (three network passes: online and
target at the n-step-shifted positions without gradient, online at the learning
positions with gradient) so that its CPU timing is a fair stand-in for the
reference's own CPU learner.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Params = Dict[str, torch.Tensor]

PARAM_SHAPES = lambda A, C=1, H=512: {  # model.py:39-63 (conv1 in-channels generalised to C)
    "feature.0.weight": (32, C, 8, 8), "feature.0.bias": (32,),
    "feature.2.weight": (64, 32, 4, 4), "feature.2.bias": (64,),
    "feature.4.weight": (64, 64, 3, 3), "feature.4.bias": (64,),
    "feature.7.weight": (512, 3136), "feature.7.bias": (512,),
    "recurrent.weight_ih_l0": (4 * H, 512 + A + 1), "recurrent.weight_hh_l0": (4 * H, H),
    "recurrent.bias_ih_l0": (4 * H,), "recurrent.bias_hh_l0": (4 * H,),
    "advantage.0.weight": (H, H), "advantage.0.bias": (H,),
    "advantage.2.weight": (A, H), "advantage.2.bias": (A,),
    "value.0.weight": (H, H), "value.0.bias": (H,),
    "value.2.weight": (1, H), "value.2.bias": (1,),
}


def init_params(action_dim: int, in_channels: int = 1, hidden_dim: int = 512, seed: int = 0,
                gain: float = 1.0) -> Params:
    """Seeded, platform-independent stand-in for PyTorch's default init.

    Uses numpy's PCG64 (stable across platforms) with the same U(-1/sqrt(fan_in),
    +1/sqrt(fan_in)) bounds PyTorch applies to Conv2d/Linear/LSTM, so tests and
    fixtures can regenerate identical weights anywhere without shipping 17 MB.
    """
    rng = np.random.default_rng(seed)
    H = hidden_dim
    out: Params = {}
    for name, shape in PARAM_SHAPES(action_dim, in_channels, H).items():
        if name.startswith("recurrent"):
            bound = 1.0 / math.sqrt(H)
        else:
            wshape = PARAM_SHAPES(action_dim, in_channels, H)[name.rsplit(".", 1)[0] + ".weight"]
            bound = 1.0 / math.sqrt(int(np.prod(wshape[1:])))
        out[name] = torch.from_numpy(
            (rng.uniform(-bound, bound, size=shape) * gain).astype(np.float32))
    return out


# --------------------------------------------------------------------------- network
def encode(p: Params, frames_u8_or_f: torch.Tensor) -> torch.Tensor:
    """model.py:39-49 on frames already divided by 255 (worker.py:342)."""
    x = F.relu(F.conv2d(frames_u8_or_f, p["feature.0.weight"], p["feature.0.bias"], stride=4))
    x = F.relu(F.conv2d(x, p["feature.2.weight"], p["feature.2.bias"], stride=2))
    x = F.relu(F.conv2d(x, p["feature.4.weight"], p["feature.4.bias"], stride=1))
    x = x.flatten(1)                                   # (c, h, w) order, nn.Flatten on NCHW
    return F.relu(F.linear(x, p["feature.7.weight"], p["feature.7.bias"]))


def lstm_unroll(p: Params, u: torch.Tensor, h0: torch.Tensor, c0: torch.Tensor,
                lengths: torch.Tensor) -> torch.Tensor:
    """Single-layer LSTM over ragged sequences (model.py:51,95-100,134-141).

    ``u`` is (B, T, 512+A+1); sequence n is advanced for t < lengths[n] only,
    exactly what ``pack_padded_sequence`` + ``nn.LSTM`` do; gate order i,f,g,o.
    Returns all hidden states (B, T, H) (rows past the length are zero, like
    ``pad_packed_sequence``).
    """
    B, T, _ = u.shape
    H = h0.shape[1]
    xproj = F.linear(u, p["recurrent.weight_ih_l0"], p["recurrent.bias_ih_l0"])
    h, c = h0, c0
    outs = []
    lengths = lengths.to(torch.int64)
    for t in range(T):
        gates = xproj[:, t] + F.linear(h, p["recurrent.weight_hh_l0"], p["recurrent.bias_hh_l0"])
        i, f, g, o = gates.split(H, dim=1)
        c_new = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
        h_new = torch.sigmoid(o) * torch.tanh(c_new)
        live = (t < lengths).unsqueeze(1)
        c = torch.where(live, c_new, c)
        h = torch.where(live, h_new, h)
        outs.append(torch.where(live, h_new, torch.zeros_like(h_new)))
    return torch.stack(outs, dim=1)


def lstm_unroll_packed(p: Params, u: torch.Tensor, h0: torch.Tensor, c0: torch.Tensor,
                       lengths: torch.Tensor) -> torch.Tensor:
    """The same recurrence through torch's packed-sequence LSTM op -- literally what the reference calls
    (pack_padded_sequence + nn.LSTM + pad_packed_sequence, model.py:95-100).  Used for the timed CPU baseline so
    that the port is as fast as the reference's own code; ``lstm_unroll`` stays the readable specification
    (tests check the two agree)."""
    from torch.nn.utils.rnn import pack_padded_sequence, pad_packed_sequence
    packed = pack_padded_sequence(u, lengths.to(torch.int64).cpu(), batch_first=True, enforce_sorted=False)
    weights = [p["recurrent.weight_ih_l0"], p["recurrent.weight_hh_l0"], p["recurrent.bias_ih_l0"], p["recurrent.bias_hh_l0"]]
    hx = (h0[packed.sorted_indices].unsqueeze(0), c0[packed.sorted_indices].unsqueeze(0))
    out, _, _ = torch._VF.lstm(packed.data, packed.batch_sizes, hx, weights, True, 1, 0.0, False, False)
    out = torch.nn.utils.rnn.PackedSequence(out, packed.batch_sizes, packed.sorted_indices, packed.unsorted_indices)
    hs, _ = pad_packed_sequence(out, batch_first=True, total_length=u.shape[1])
    return hs


_LSTM_IMPL = {"loop": lstm_unroll, "packed": lstm_unroll_packed}
LSTM_MODE = "loop"


def dueling_head(p: Params, hidden_rows: torch.Tensor) -> torch.Tensor:
    """model.py:115-117 / 145-148."""
    adv = F.linear(F.relu(F.linear(hidden_rows, p["advantage.0.weight"], p["advantage.0.bias"])),
                   p["advantage.2.weight"], p["advantage.2.bias"])
    val = F.linear(F.relu(F.linear(hidden_rows, p["value.0.weight"], p["value.0.bias"])),
                   p["value.2.weight"], p["value.2.bias"])
    return val + adv - adv.mean(1, keepdim=True)


def _recurrent_input(p: Params, obs: torch.Tensor, last_action: torch.Tensor,
                     last_reward: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    B, T = obs.shape[:2]
    latent = encode(p, obs.reshape(B * T, *obs.shape[2:]))           # model.py:86-88 / 126-130
    u = torch.cat((latent, last_action.reshape(B * T, -1), last_reward.reshape(B * T, 1)), dim=1)
    return u.view(B, T, -1), latent.view(B, T, -1)


def shifted_rows(burn_in, learning, forward, max_forward: int = 5):
    """Hidden-state index of every output row of ``calculate_q_`` (model.py:102-111).

    Row t of sequence n reads h[min(b+F+t, b+l+f-1)] with F = max_forward: the
    slice [b+F : b+l+f] followed by min(F-f, l) repeats of the last state.
    Returned sequence-major as (seq_index, time_index) int64 arrays.
    """
    seq, tim = [], []
    for n, (b, l, f) in enumerate(zip(burn_in.tolist(), learning.tolist(), forward.tolist())):
        for t in range(l):
            seq.append(n)
            tim.append(min(b + max_forward + t, b + l + f - 1))
    return torch.tensor(seq, dtype=torch.int64), torch.tensor(tim, dtype=torch.int64)


def learning_rows(burn_in, learning):
    """Rows of ``calculate_q``: h[b .. b+l-1] per sequence (model.py:143)."""
    seq, tim = [], []
    for n, (b, l) in enumerate(zip(burn_in.tolist(), learning.tolist())):
        for t in range(l):
            seq.append(n)
            tim.append(b + t)
    return torch.tensor(seq, dtype=torch.int64), torch.tensor(tim, dtype=torch.int64)


def calculate_q_shifted(p: Params, obs, last_action, last_reward, h0, c0, burn_in, learning, forward,
                        max_forward: int = 5, want_hidden: bool = False):
    """``Network.calculate_q_`` (model.py:81-119).  obs is float, already /255."""
    u, latent = _recurrent_input(p, obs, last_action, last_reward)
    lengths = burn_in.to(torch.int64) + learning.to(torch.int64) + forward.to(torch.int64)
    hs = _LSTM_IMPL[LSTM_MODE](p, u, h0, c0, lengths)
    seq, tim = shifted_rows(burn_in, learning, forward, max_forward)
    q = dueling_head(p, hs[seq, tim])
    return (q, hs, latent) if want_hidden else q


def calculate_q(p: Params, obs, last_action, last_reward, h0, c0, burn_in, learning,
                want_hidden: bool = False):
    """``Network.calculate_q`` (model.py:122-150)."""
    u, latent = _recurrent_input(p, obs, last_action, last_reward)
    lengths = burn_in.to(torch.int64) + learning.to(torch.int64)
    hs = _LSTM_IMPL[LSTM_MODE](p, u, h0, c0, lengths)
    seq, tim = learning_rows(burn_in, learning)
    q = dueling_head(p, hs[seq, tim])
    return (q, hs, latent) if want_hidden else q


# --------------------------------------------------------------------------- TD math
def value_rescale(x: torch.Tensor, eps: float = 1e-3) -> torch.Tensor:
    """h(x), worker.py:383-385."""
    return x.sign() * ((x.abs() + 1).sqrt() - 1) + eps * x


def inverse_value_rescale(x: torch.Tensor, eps: float = 1e-3) -> torch.Tensor:
    """h^-1(x), worker.py:387-390."""
    t = ((1 + 4 * eps * (x.abs() + 1 + eps)).sqrt() - 1) / (2 * eps)
    return x.sign() * (t.square() - 1)


def td_target_ieee(q_tgt: np.ndarray, R: np.ndarray, G: np.ndarray) -> np.ndarray:
    """h(R + G*h^-1(q)) with the reference's float32 operation sequence, every op IEEE
    round-to-nearest (NumPy).  torch's CPU ``sqrt`` is not correctly rounded for ~1.3 % of
    inputs and h^-1 amplifies that single ulp to ~1e-4 when |q| >= 10, so this -- not the torch
    evaluation -- is the tight check for the CUDA kernel (which uses IEEE sqrt, as torch's own
    CUDA path does)."""
    f = np.float32
    x = q_tgt.astype(np.float32)
    a = (np.abs(x) + f(1)) + f(1e-3)
    s = a * f(4 * 1e-3) + f(1)
    t = (np.sqrt(s) - f(1)) / f(2 * 1e-3)
    inv = np.sign(x) * (t * t - f(1))
    y = R.astype(np.float32) + G.astype(np.float32) * inv
    return np.sign(y) * (np.sqrt(np.abs(y) + f(1)) - f(1)) + f(1e-3) * y


def mixed_priorities(td: np.ndarray, learning_steps: np.ndarray) -> np.ndarray:
    """``calculate_mixed_td_errors`` (worker.py:268-276) with the NumPy-1.x
    accumulator semantics the reference was written for (int64 running offset)."""
    out = np.empty(learning_steps.shape, dtype=td.dtype)
    start = 0
    for n, steps in enumerate(learning_steps.astype(np.int64)):
        seg = td[start:start + steps]
        out[n] = 0.9 * seg.max() + 0.1 * seg.mean()
        start += int(steps)
    return out


# --------------------------------------------------------------------------- one update
@dataclass
class LearnerState:
    online: Params
    target: Params
    lr: float = 1e-4
    adam_eps: float = 1e-3
    betas: Tuple[float, float] = (0.9, 0.999)
    grad_norm: float = 40.0
    step: int = 0
    exp_avg: Optional[Params] = None
    exp_avg_sq: Optional[Params] = None

    def __post_init__(self):
        if self.exp_avg is None:
            self.exp_avg = {k: torch.zeros_like(v) for k, v in self.online.items()}
            self.exp_avg_sq = {k: torch.zeros_like(v) for k, v in self.online.items()}


@dataclass
class Batch:
    """The learner-visible part of the 14-tuple of worker.py:219-238."""
    obs: torch.Tensor            # u8 (B, T, C, 84, 84)
    last_action: torch.Tensor    # bool/float (B, T, A)
    last_reward: torch.Tensor    # f32 (B, T)
    hidden: torch.Tensor         # f32 (2, B, H)   [0]=h0 [1]=c0
    action: torch.Tensor         # u8/int (sumL, 1)
    n_step_reward: torch.Tensor  # f32 (sumL,)
    n_step_gamma: torch.Tensor   # f32 (sumL,)
    burn_in: torch.Tensor        # u8 (B,)
    learning: torch.Tensor       # u8 (B,)
    forward: torch.Tensor        # u8 (B,)
    is_weights: torch.Tensor     # f32 (sumL,)

    @staticmethod
    def from_tuple(t) -> "Batch":
        return Batch(t[0], t[1], t[2], t[3], t[4], t[5], t[6], t[7], t[8], t[9], t[11])


def td_and_loss(q, qn_online, qn_target, action, R, G, is_w):
    """worker.py:346-357: double-Q selection, rescaled n-step target, IS-weighted MSE."""
    a_star = qn_online.argmax(1, keepdim=True)
    q_tgt = qn_target.gather(1, a_star).squeeze(1)
    target = value_rescale(R + G * inverse_value_rescale(q_tgt))
    q_a = q.gather(1, action.long().view(-1, 1)).squeeze(1)
    loss = (is_w * (q_a - target) ** 2).mean()
    td = (target - q_a).detach().abs()
    return loss, td, target, q_a


def learner_update(st: LearnerState, batch: Batch, max_forward: int = 5, apply: bool = True) -> dict:
    """One iteration of ``Learner.run`` (worker.py:330-365) minus queues/IO.

    Returns every intermediate the GPU tests compare against.
    """
    obs = batch.obs.float() / 255                                   # worker.py:336,342
    la = batch.last_action.float()
    lr_ = batch.last_reward.float()
    h0, c0 = batch.hidden[0].contiguous(), batch.hidden[1].contiguous()   # worker.py:340

    with torch.no_grad():                                            # worker.py:345-347
        qn_on = calculate_q_shifted(st.online, obs, la, lr_, h0, c0, batch.burn_in, batch.learning,
                                    batch.forward, max_forward)
        qn_tg = calculate_q_shifted(st.target, obs, la, lr_, h0, c0, batch.burn_in, batch.learning,
                                    batch.forward, max_forward)

    online = {k: v.detach().clone().requires_grad_(True) for k, v in st.online.items()}
    q, hs, latent = calculate_q(online, obs, la, lr_, h0, c0, batch.burn_in, batch.learning,
                                want_hidden=True)                    # worker.py:352
    loss, td, target, q_a = td_and_loss(q, qn_on, qn_tg, batch.action, batch.n_step_reward,
                                        batch.n_step_gamma, batch.is_weights)
    prio = mixed_priorities(td.numpy().astype(np.float32), batch.learning.numpy())  # worker.py:357-359

    loss.backward()                                                  # worker.py:362-363
    grads = {k: v.grad.detach() for k, v in online.items()}
    total_norm = torch.sqrt(sum((g.double() ** 2).sum() for g in grads.values())).float()
    out = dict(loss=float(loss.item()), td=td.numpy(), priorities=prio, q=q.detach(), qn_online=qn_on,
               qn_target=qn_tg, target=target.detach(), q_a=q_a.detach(), hidden=hs.detach(),
               latent=latent.detach(), grads=grads, grad_norm=float(total_norm))
    if apply:
        adam_step(st, grads)
    return out


def adam_step(st: LearnerState, grads: Params) -> None:
    """``clip_grad_norm_(40)`` then ``Adam(lr, eps)`` (worker.py:289,364-365)."""
    total_norm = torch.sqrt(sum((g.float() ** 2).sum() for g in grads.values()))
    coef = torch.clamp(st.grad_norm / (total_norm + 1e-6), max=1.0)
    st.step += 1
    b1, b2 = st.betas
    bc1 = 1 - b1 ** st.step
    bc2 = 1 - b2 ** st.step
    for k, p in st.online.items():
        g = grads[k] * coef
        st.exp_avg[k].mul_(b1).add_(g, alpha=1 - b1)
        st.exp_avg_sq[k].mul_(b2).addcmul_(g, g, value=1 - b2)
        denom = (st.exp_avg_sq[k].sqrt() / math.sqrt(bc2)).add_(st.adam_eps)
        p.addcdiv_(st.exp_avg[k], denom, value=-st.lr / bc1)
