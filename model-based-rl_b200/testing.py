"""Deterministic, batch-size-invariant stand-in network used for bit-exact search parity.

The reference drives its network one leaf at a time (mcts.py:96) while the engine drives it for all
games at once; real GEMMs are not bit-identical across batch sizes, so search parity is checked
with a network whose outputs are a pure integer function of (hidden_state, action)
(SURVEY.md section 4 item 2).  The same function is restated in oracle/mz_oracle.c
(orc_hashnet_next / orc_hashnet_outputs).

It implements the reference's network interface (networks.py:9, 26-34): initial_inference(obs) and
recurrent_inference(hidden_state, action) -> NetworkOutput(value, reward, policy_logits,
hidden_state).  The hidden state is one int64 per row.
"""
from collections import namedtuple

import torch

NetworkOutput = namedtuple('network_output', ('value', 'reward', 'policy_logits', 'hidden_state'))

_M64 = (1 << 64) - 1


def _s64(c):
  c &= _M64
  return c - (1 << 64) if c >= (1 << 63) else c


_GOLDEN = _s64(0x9E3779B97F4A7C15)
_MUL1 = _s64(0xBF58476D1CE4E5B9)
_MUL2 = _s64(0x94D049BB133111EB)
_KACT = _s64(0xD6E8FEB86659FD93)
_KOUT = 0xA0761D6478BD642F


def _lsr(z, k):
  return (z >> k) & ((1 << (64 - k)) - 1)


def _splitmix64(x):
  z = x + _GOLDEN
  z = (z ^ _lsr(z, 30)) * _MUL1
  z = (z ^ _lsr(z, 27)) * _MUL2
  return z ^ _lsr(z, 31)


def _unit(o):
  u = _lsr(o, 40).to(torch.float32) * (2.0 ** -24)
  return 2.0 * u - 1.0


class HashNetwork(object):
  """value/reward/logits are exact float32 functions of a 64-bit state; works on any device."""

  accepts_device_actions = True  # the engine may pass `action` as an int tensor on the device
  training = False

  def __init__(self, action_space, value_scale=1.0, reward_scale=0.5, logit_scale=2.0,
               reward_density=3, device='cpu'):
    self.action_space = action_space
    self.value_scale = float(value_scale)
    self.reward_scale = float(reward_scale)
    self.logit_scale = float(logit_scale)
    self.reward_density = int(reward_density)
    self.device = torch.device(device)

  def to(self, device):
    self.device = torch.device(device)
    return self

  def eval(self):
    return self

  def _outputs(self, state):
    o0 = _splitmix64(state + _s64(1 * _KOUT))
    o1 = _splitmix64(state + _s64(2 * _KOUT))
    value = _unit(o0) * self.value_scale
    reward = torch.where((o1 & 7) < self.reward_density, _unit(o1) * self.reward_scale,
                         torch.zeros((), dtype=torch.float32, device=state.device))
    cols = []
    for a in range(self.action_space):
      cols.append(_unit(_splitmix64(state + _s64((a + 3) * _KOUT))))
    logits = torch.cat(cols, dim=1) * self.logit_scale
    return value, reward, logits

  def initial_inference(self, observation):
    """`observation` is the int64 root state, shape [B, 1]."""
    state = observation.to(self.device).to(torch.int64).view(-1, 1)
    value, _, logits = self._outputs(state)
    return NetworkOutput(value, 0, logits, state)

  def recurrent_inference(self, hidden_state, action):
    state = hidden_state.to(torch.int64).view(-1, 1)
    a = torch.as_tensor(action, device=state.device).to(torch.int64).view(-1, 1)
    nxt = _splitmix64(state ^ ((a + 1) * _KACT))
    value, reward, logits = self._outputs(nxt)
    return NetworkOutput(value, reward, logits, nxt)


class ObsHashNetwork(HashNetwork):
  """HashNetwork whose root state is a hash of the observation row (for driver-level parity tests):
  state = sum_i (obs_i + 2) * 5^i over the first 24 entries, exact in int64."""

  def initial_inference(self, observation):
    obs = torch.as_tensor(observation).to(self.device).reshape(observation.shape[0], -1)
    n = min(obs.shape[1], 24)
    w = torch.tensor([5 ** i for i in range(n)], dtype=torch.int64, device=obs.device)
    state = ((obs[:, :n].to(torch.int64) + 2) * w).sum(dim=1, keepdim=True)
    return super().initial_inference(state)
