"""Multi-GPU plumbing of the hot path (one process per GPU, torch.distributed).

The search itself needs no communication: every game's tree, MinMaxStats and trajectory are
independent (one MCTS + Game per actor in the reference, actors.py:29-31), so rank r simply owns a
contiguous slice of the games.  The only collectives are the two the reference performs through
Ray's object store: the learner -> actors weight hand-off
(learners.py:85-86, actors.py:81-85) and, when the learner is data parallel, the gradient
all-reduce.  Backend: "nccl" on GPUs (NVLink / NVSwitch), "gloo" in the CPU tests.
"""
import torch
import torch.distributed as dist


def world():
  if dist.is_available() and dist.is_initialized():
    return dist.get_rank(), dist.get_world_size()
  return 0, 1


def game_slice(total_games, rank=None, world_size=None):
  """[lo, hi) of the games owned by `rank` when `total_games` are sharded as evenly as possible."""
  r, w = world()
  rank = r if rank is None else rank
  world_size = w if world_size is None else world_size
  if not 0 <= rank < world_size:
    raise ValueError("rank %d outside world of %d" % (rank, world_size))
  return (total_games * rank) // world_size, (total_games * (rank + 1)) // world_size


def max_over_ranks(value, device=None):
  """Max of a host scalar over all ranks (timing: a step takes as long as its slowest rank)."""
  _, w = world()
  if w == 1:
    return float(value)
  t = torch.tensor([float(value)], dtype=torch.float64, device=device)
  dist.all_reduce(t, op=dist.ReduceOp.MAX)
  return float(t.item())


def broadcast_weights(state_dict, src=0):
  """Learner -> self-play ranks weight hand-off: every tensor of `state_dict` is overwritten in place
  with rank `src`'s values (one flat buffer per dtype, so the FCNetwork is a single small message)."""
  _, w = world()
  if w == 1:
    return state_dict
  by_dtype = {}
  for k in sorted(state_dict):
    by_dtype.setdefault(state_dict[k].dtype, []).append(k)
  for dtype, keys in by_dtype.items():
    flat = torch.cat([state_dict[k].reshape(-1) for k in keys])
    dist.broadcast(flat, src=src)
    off = 0
    for k in keys:
      n = state_dict[k].numel()
      state_dict[k].copy_(flat[off:off + n].view_as(state_dict[k]))
      off += n
  return state_dict


def allreduce_gradients(parameters, average=True):
  """Data-parallel learner step: sums (or averages) .grad over ranks through one flat buffer."""
  _, w = world()
  grads = [p.grad for p in parameters if p.grad is not None]
  if w == 1 or not grads:
    return
  flat = torch.cat([g.reshape(-1) for g in grads])
  dist.all_reduce(flat, op=dist.ReduceOp.SUM)
  if average:
    flat /= w
  off = 0
  for g in grads:
    n = g.numel()
    g.copy_(flat[off:off + n].view_as(g))
    off += n
