"""World-size-2 gloo tests (CPU) of the N>1 host logic: game sharding, max-over-ranks timing,
weight broadcast and gradient all-reduce (model-based-rl_b200/parallel.py)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
  s = socket.socket()
  s.bind(("127.0.0.1", 0))
  port = s.getsockname()[1]
  s.close()
  return port


def _worker(rank, world_size, port, out):
  os.environ["MASTER_ADDR"] = "127.0.0.1"
  os.environ["MASTER_PORT"] = str(port)
  dist.init_process_group("gloo", rank=rank, world_size=world_size)
  from model_based_rl_b200 import parallel
  try:
    lo, hi = parallel.game_slice(4097)
    slices = [None] * world_size
    dist.all_gather_object(slices, (lo, hi))
    assert slices[0][0] == 0 and slices[-1][1] == 4097
    assert all(slices[i][1] == slices[i + 1][0] for i in range(world_size - 1))
    assert parallel.max_over_ranks(1.0 + rank) == float(world_size)
    torch.manual_seed(rank)
    sd = {"a.weight": torch.randn(3, 5), "a.bias": torch.randn(5), "steps": torch.tensor([rank], dtype=torch.int64)}
    want = None
    if rank == 0:
      want = {k: v.clone() for k, v in sd.items()}
    box = [want]
    dist.broadcast_object_list(box, src=0)
    parallel.broadcast_weights(sd, src=0)
    for k in sd:
      assert torch.equal(sd[k], box[0][k]), k
    lin = torch.nn.Linear(4, 2)
    for p in lin.parameters():
      p.grad = torch.full_like(p, float(rank + 1))
    parallel.allreduce_gradients(list(lin.parameters()))
    mean = sum(range(1, world_size + 1)) / world_size
    for p in lin.parameters():
      assert torch.allclose(p.grad, torch.full_like(p, mean))
    out.put((rank, "ok"))
  except Exception as e:  # pragma: no cover
    out.put((rank, repr(e)))
  finally:
    dist.destroy_process_group()


def test_two_rank_gloo():
  ctx = mp.get_context("spawn")
  out = ctx.Queue()
  port = _free_port()
  procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
  for p in procs:
    p.start()
  res = [out.get(timeout=120) for _ in procs]
  for p in procs:
    p.join(timeout=60)
  assert sorted(res) == [(0, "ok"), (1, "ok")], res
