"""The slice of the reference's Config (config.py:7-81) that the hot path uses, on CUDA kernels.

`Config(args_dict)` derives the same support fields as the reference and offers the same methods:
inverse_value_transform / inverse_reward_transform / inverse_transform ("support_to_scalar"),
value_phi / reward_phi / scalar_to_support (two-hot projection), scalar_transform (h(x)),
visit_softmax_temperature and select_action.  Tensor methods take and return CUDA tensors.
"""
import numpy as np
import torch

from . import _lib


def _f32(x):
  if not x.is_cuda:
    _lib.require_cuda()
    raise RuntimeError("the B200 transforms take CUDA tensors (there is no CPU fallback)")
  return x.to(torch.float32).contiguous()


class Config(object):

  def __init__(self, args):
    self.__dict__.update(args)
    self.value_support_min, self.value_support_max = self.value_support
    self.reward_support_min, self.reward_support_max = self.reward_support
    self.value_support_range = list(range(self.value_support_min, self.value_support_max + 1))
    self.value_support_size = len(self.value_support_range)
    self.reward_support_range = list(range(self.reward_support_min, self.reward_support_max + 1))
    self.reward_support_size = len(self.reward_support_range)

  # -- support -> scalar (config.py:21-33) -------------------------------------------------------
  def inverse_reward_transform(self, reward_logits):
    return self.inverse_transform(reward_logits, self.reward_support_range)

  def inverse_value_transform(self, value_logits):
    return self.inverse_transform(value_logits, self.value_support_range)

  def inverse_transform(self, logits, scalar_support_range):
    logits = _f32(logits)
    n, bins = logits.shape
    mn, mx = int(scalar_support_range[0]), int(scalar_support_range[-1])
    if bins != mx - mn + 1:
      raise ValueError("logits have %d bins, support has %d" % (bins, mx - mn + 1))
    out = torch.empty((n, 1), dtype=torch.float32, device=logits.device)
    _lib.check(_lib.load().mz_support_to_scalar(n, _lib.ptr(logits), mn, mx,
                                                int(bool(getattr(self, 'no_target_transform', False))),
                                                _lib.ptr(out), _lib.current_stream()),
               "mz_support_to_scalar")
    return out

  # -- scalar -> support (config.py:35-39, 56-68) ------------------------------------------------
  def value_phi(self, x):
    return self.scalar_to_support(x, self.value_support_min, self.value_support_max,
                                  self.value_support_size)

  def reward_phi(self, x):
    return self.scalar_to_support(x, self.reward_support_min, self.reward_support_max,
                                  self.reward_support_size)

  @staticmethod
  def scalar_to_support(x, min, max, support_size):
    """x [B, K] float32 CUDA tensor, clamped IN PLACE like the reference (x.clamp_)."""
    if support_size != max - min + 1:
      raise ValueError("support_size must be max - min + 1")
    if not (x.is_cuda and x.dtype == torch.float32 and x.is_contiguous()):
      raise RuntimeError("scalar_to_support needs a contiguous float32 CUDA tensor (clamped in place)")
    support = torch.empty(tuple(x.shape) + (support_size,), dtype=torch.float32, device=x.device)
    _lib.check(_lib.load().mz_scalar_to_support(x.numel(), _lib.ptr(x), int(min), int(max), 1,
                                                _lib.ptr(support), _lib.current_stream()),
               "mz_scalar_to_support")
    return support

  @staticmethod
  def scalar_transform(x):
    x = _f32(x)
    out = torch.empty_like(x)
    _lib.check(_lib.load().mz_scalar_transform(x.numel(), _lib.ptr(x), _lib.ptr(out),
                                               _lib.current_stream()), "mz_scalar_transform")
    return out

  # -- self-play helpers (config.py:41-49, 70-81) ------------------------------------------------
  def visit_softmax_temperature(self, training_step):
    step1, step2 = self.visit_softmax_steps
    temp1, temp2, temp3 = self.visit_softmax_temperatures
    if training_step <= step1:
      return temp1
    if training_step <= step2:
      return temp2
    return temp3

  @staticmethod
  def select_action(node, temperature=0.):
    """Host form for a single Node (B=1 drop-in).  The batched device form is
    BatchedMCTS.select_action, which takes the uniforms from the caller."""
    actions = list(node.children.keys())
    counts = np.array([child.visit_count for child in node.children.values()])
    if temperature:
      weights = counts ** (1 / temperature)
      idx = np.random.choice(len(actions), p=weights / weights.sum())
    else:
      idx = np.random.choice(np.where(counts == counts.max())[0])
    return actions[idx]
