"""ORACLE (test infrastructure, never on the product path): functional float32 restatement of the
reference's MuZeroNetwork (networks.py:372-554) on a plain state dict with the reference's keys.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline may import this.

Pinned to tests/golden/muzero_net.npz, which holds outputs of the unmodified reference class
(eval mode) for weights drawn by `seeded_state_dict` below (tests/test_oracle_golden.py)."""
import math

import torch
import torch.nn.functional as F

BN_EPS = 1e-5


def _bn(x, sd, p):  # nn.BatchNorm2d in eval mode
  return F.batch_norm(x, sd[p + '.running_mean'], sd[p + '.running_var'], sd[p + '.weight'],
                      sd[p + '.bias'], False, 0.0, BN_EPS)


def resblock(x, sd, p):  # ResidualBlock.forward networks.py:384-391
  out = F.relu(_bn(F.conv2d(x, sd[p + '.conv1.weight'], None, 1, 1), sd, p + '.bn1'))
  out = _bn(F.conv2d(out, sd[p + '.conv2.weight'], None, 1, 1), sd, p + '.bn2')
  return F.relu(out + x)


def scale_state(s):  # networks.py:543-547
  mn = s.min(dim=1, keepdim=True)[0]
  mx = s.max(dim=1, keepdim=True)[0]
  return (s - mn) / (mx - mn)


def representation(obs, sd):  # MuZeroRepresentation.forward networks.py:412-426 + scale_state
  p = 'representation_head.'
  out = F.conv2d(obs, sd[p + 'conv1.weight'], sd[p + 'conv1.bias'], 2, 1)
  for i in range(2):
    out = resblock(out, sd, p + 'resblocks1.%d' % i)
  out = F.conv2d(out, sd[p + 'conv2.weight'], sd[p + 'conv2.bias'], 2, 1)
  for i in range(3):
    out = resblock(out, sd, p + 'resblocks2.%d' % i)
  out = F.avg_pool2d(out, 3, 2, 1)
  for i in range(3):
    out = resblock(out, sd, p + 'resblocks3.%d' % i)
  out = F.avg_pool2d(out, 3, 2, 1)
  for i in range(16):
    out = resblock(out, sd, p + 'resblocks.%d' % i)
  return scale_state(out)


def dynamics_logits(hidden, actions, sd, action_space):
  """MuZeroNetwork.attach_action + MuZeroDynamics.forward (networks.py:431-449, 536-541):
  -> (scaled next state, reward support logits)."""
  p = 'dynamics_head.'
  a = torch.as_tensor(actions, dtype=torch.float32, device=hidden.device)
  plane = (a / action_space)[:, None, None, None].expand(-1, 1, hidden.shape[2], hidden.shape[3])
  out = F.conv2d(torch.cat((hidden, plane), dim=1), sd[p + 'conv.weight'], sd[p + 'conv.bias'], 1, 1)
  out = F.relu(_bn(out, sd, p + 'bn'))
  for i in range(16):
    out = resblock(out, sd, p + 'resblocks.%d' % i)
  r = F.relu(F.linear(out.reshape(out.shape[0], -1), sd[p + 'fc1.weight'], sd[p + 'fc1.bias']))
  r = F.linear(r, sd[p + 'fc2.weight'], sd[p + 'fc2.bias'])
  return scale_state(out), r


def prediction_logits(hidden, sd):  # MuZeroPrediction.forward networks.py:465-480
  p = 'prediction_head.'
  out = hidden
  for i in range(16):
    out = resblock(out, sd, p + 'resblocks.%d' % i)
  flat = out.reshape(out.shape[0], -1)
  v = F.linear(F.relu(F.linear(flat, sd[p + 'fc_value.weight'], sd[p + 'fc_value.bias'])),
               sd[p + 'fc_value_o.weight'], sd[p + 'fc_value_o.bias'])
  pol = F.linear(F.relu(F.linear(flat, sd[p + 'fc_policy.weight'], sd[p + 'fc_policy.bias'])),
                 sd[p + 'fc_policy_o.weight'], sd[p + 'fc_policy_o.bias'])
  return pol, v


def inverse_transform(logits, support_min, support_max, no_target_transform=False):  # config.py:27-33
  support = torch.arange(support_min, support_max + 1, dtype=torch.float32, device=logits.device)
  x = (torch.softmax(logits, dim=1) * support).sum(dim=1, keepdim=True)
  if not no_target_transform:
    eps = 0.001
    x = torch.sign(x) * (((torch.sqrt(1 + 4 * eps * (torch.abs(x) + 1 + eps)) - 1) / (2 * eps)) ** 2 - 1)
  return x


def initial_inference(obs, sd, support=(-15, 15)):  # networks.py:26-29 in eval mode
  h = representation(obs, sd)
  pol, v = prediction_logits(h, sd)
  return inverse_transform(v, *support), pol, h


def recurrent_inference(hidden, actions, sd, action_space, support=(-15, 15)):  # networks.py:31-34
  h, r = dynamics_logits(hidden, actions, sd, action_space)
  pol, v = prediction_logits(h, sd)
  return inverse_transform(v, *support), inverse_transform(r, *support), pol, h


def state_dict_shapes(input_channels, action_space, value_bins=31, reward_bins=31):
  """Every tensor of MuZeroNetwork.state_dict() in registration order (networks.py:397-410,
  433-438, 456-463, 493-496)."""
  shapes = []

  def conv(p, cin, cout, bias):
    shapes.append((p + '.weight', (cout, cin, 3, 3)))
    if bias:
      shapes.append((p + '.bias', (cout,)))

  def bn(p, c):
    shapes.extend([(p + '.weight', (c,)), (p + '.bias', (c,)), (p + '.running_mean', (c,)),
                   (p + '.running_var', (c,)), (p + '.num_batches_tracked', ())])

  def blocks(p, n, c):
    for i in range(n):
      b = '%s.%d' % (p, i)
      conv(b + '.conv1', c, c, False)
      bn(b + '.bn1', c)
      conv(b + '.conv2', c, c, False)
      bn(b + '.bn2', c)

  def linear(p, i, o):
    shapes.extend([(p + '.weight', (o, i)), (p + '.bias', (o,))])

  r = 'representation_head'
  conv(r + '.conv1', input_channels, 64, True)
  blocks(r + '.resblocks1', 2, 64)
  conv(r + '.conv2', 64, 128, True)
  blocks(r + '.resblocks2', 3, 128)
  blocks(r + '.resblocks3', 3, 128)
  blocks(r + '.resblocks', 16, 128)
  q = 'prediction_head'
  blocks(q + '.resblocks', 16, 128)
  linear(q + '.fc_value', 4608, 512)
  linear(q + '.fc_value_o', 512, value_bins)
  linear(q + '.fc_policy', 4608, 512)
  linear(q + '.fc_policy_o', 512, action_space)
  d = 'dynamics_head'
  conv(d + '.conv', 129, 128, True)
  bn(d + '.bn', 128)
  blocks(d + '.resblocks', 16, 128)
  linear(d + '.fc1', 4608, 512)
  linear(d + '.fc2', 512, reward_bins)
  return shapes


def seeded_state_dict(input_channels, action_space, seed, value_bins=31, reward_bins=31):
  """Deterministic weights by tensor name order (independent of module construction): conv /
  linear weights ~ U(+-1/sqrt(fan_in)) like torch's default, BatchNorm affine and running
  statistics perturbed so that folding them is actually exercised."""
  g = torch.Generator().manual_seed(seed)
  sd = {}
  for name, shape in state_dict_shapes(input_channels, action_space, value_bins, reward_bins):
    if name.endswith('num_batches_tracked'):
      sd[name] = torch.tensor(7, dtype=torch.int64)
    elif name.endswith('running_var'):
      sd[name] = 0.5 + torch.rand(shape, generator=g)
    elif name.endswith('running_mean'):
      sd[name] = 0.2 * torch.randn(shape, generator=g)
    elif '.bn' in name and name.endswith('.weight'):
      sd[name] = 0.6 + 0.5 * torch.rand(shape, generator=g)
    elif '.bn' in name and name.endswith('.bias'):
      sd[name] = 0.1 * torch.randn(shape, generator=g)
    elif name.endswith('.weight'):
      fan_in = int(torch.tensor(shape[1:]).prod())
      # gain sqrt(3) keeps activations O(1) through 16 residual blocks with the BN statistics above
      bound = math.sqrt(3.0 / fan_in)
      sd[name] = (torch.rand(shape, generator=g) * 2 - 1) * bound
    else:
      sd[name] = (torch.rand(shape, generator=g) * 2 - 1) * 0.05
  return sd
