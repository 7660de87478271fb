"""torch (CPU, float32) restatement of the reference's FCNetwork inference -- TEST INFRASTRUCTURE /
CPU BASELINE only.  Same layer shapes and state-dict keys as networks.py:55-174, eval-mode outputs
(scalars after Config.inverse_transform, config.py:27-33).  Pinned against golden outputs of the
unmodified reference module in tests/test_oracle_golden.py."""
from collections import namedtuple

import torch
import torch.nn as nn
import torch.nn.functional as F

NetworkOutput = namedtuple('network_output', ('value', 'reward', 'policy_logits', 'hidden_state'))

HIDDEN, WIDTH = 50, 512


class _Head(nn.Module):
  """Linear(d_in, 512) + ReLU + Linear(512, d_out); the second layer's attribute name varies."""

  def __init__(self, d_in, d_out, out_name):
    super().__init__()
    self.fc1 = nn.Linear(d_in, WIDTH)
    setattr(self, out_name, nn.Linear(WIDTH, d_out))
    self._out_name = out_name

  def forward(self, x):
    return getattr(self, self._out_name)(F.relu(self.fc1(x)))


def support_to_scalar(logits, support_min, support_max, no_target_transform=False):
  p = torch.softmax(logits, dim=1)
  support = torch.arange(support_min, support_max + 1, dtype=torch.float32).expand(p.shape)
  v = torch.sum(support * p, dim=1, keepdim=True)
  if not no_target_transform:
    v = torch.sign(v) * (((torch.sqrt(1 + 4 * 0.001 * (torch.abs(v) + 1 + 0.001)) - 1) / (2 * 0.001)) ** 2 - 1)
  return v


class FCNetworkRef(nn.Module):

  def __init__(self, input_dim, action_space, value_support=(-15, 15), reward_support=(-15, 15),
               no_target_transform=False):
    super().__init__()
    self.action_space = action_space
    self.vs, self.rs, self.no_tt = tuple(value_support), tuple(reward_support), no_target_transform
    self.representation_head = _Head(input_dim, HIDDEN, 'out')
    self.value_head = _Head(HIDDEN, self.vs[1] - self.vs[0] + 1, 'value')
    self.policy_head = _Head(HIDDEN, action_space, 'policy')
    self.reward_head = _Head(HIDDEN + action_space, self.rs[1] - self.rs[0] + 1, 'reward')
    self.transition_head = _Head(HIDDEN + action_space, HIDDEN, 'out')
    self.LN = nn.LayerNorm([HIDDEN])
    self.eval()

  def _predict(self, h):
    value = self.value_head(h)
    if not self.training:  # train mode keeps the support logits (networks.py:151-154)
      value = support_to_scalar(value, *self.vs, self.no_tt)
    return self.policy_head(h), value

  def initial_inference(self, observation):
    h = F.relu(self.LN(self.representation_head(observation.reshape(observation.shape[0], -1))))
    logits, value = self._predict(h)
    return NetworkOutput(value, 0, logits, h)

  def recurrent_inference(self, hidden_state, action):
    a = torch.as_tensor(action, dtype=torch.int64).reshape(-1, 1)
    onehot = torch.zeros((a.shape[0], self.action_space), dtype=torch.float32).scatter_(1, a, 1.0)
    x = torch.cat((hidden_state, onehot), dim=1)
    reward = self.reward_head(x)
    if not self.training:
      reward = support_to_scalar(reward, *self.rs, self.no_tt)
    h = F.relu(self.LN(self.transition_head(x)))
    logits, value = self._predict(h)
    return NetworkOutput(value, reward, logits, h)


def random_state_dict(input_dim, action_space, seed=1234):
  """Random-init weights of the FCNetwork architecture with the reference's keys (torch default
  nn.Linear / LayerNorm initialisation)."""
  g = torch.random.get_rng_state()
  torch.manual_seed(seed)
  net = FCNetworkRef(input_dim, action_space)
  torch.random.set_rng_state(g)
  return {k: v.detach().clone() for k, v in net.state_dict().items()}
