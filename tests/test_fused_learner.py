"""The learner step on the library's own forward / backward kernels (fused_learner.py, csrc/mz_learner.cu) against
the reference goldens (tests/golden/learner_*.npz: losses and weights after two steps of the reference's FCNetwork,
Config and optimisers) and against torch autograd on the same weights."""
import types

import numpy as np
import pytest
import torch

import helpers
from test_learner import CASES, _batch, _config, _sub, _weights

pytestmark = pytest.mark.gpu


def _nets(g, cfg):
  from model_based_rl_b200 import fused_learner, learners
  D, A = int(g["obs_dim"]), int(g["action_space"])
  net = fused_learner.FusedFCNetwork(D, A, "cuda", cfg)
  net.load_weights(_weights(g))
  ref = learners.FCNetworkTrain(D, A, "cuda", cfg)
  ref.load_weights(_weights(g))
  return net, ref


@pytest.mark.parametrize("graph", [False, True], ids=["eager", "cuda_graph"])
@pytest.mark.parametrize("case", CASES)
def test_fused_learner_two_steps_match_reference_golden(case, graph):
  """Same bars as the torch-module learner (tests/test_learner.py): 1e-4 relative on the losses, 2e-5 absolute on
  every weight after each of two steps -- RMSprop (breakout), AdamW with gradient clipping (ttt: the library's own
  Adam kernel), SGD without target transform (lunar_raw)."""
  from model_based_rl_b200 import fused_learner
  g = helpers.load("learner_" + case)
  cfg = _config(g)
  net, _ = _nets(g, cfg)
  learner = fused_learner.FusedLearner(cfg, net, use_graph=graph, precision="f32")
  for step in range(2):
    losses = learner.update_weights(_batch(g, step))
    np.testing.assert_allclose(losses.cpu().numpy(), g["s%d_losses" % step], rtol=1e-4)
    np.testing.assert_allclose(learner.last_errors.cpu().numpy(), g["s%d_new_errors" % step], rtol=0, atol=2e-3)
    for k, v in net.state_dict().items():
      np.testing.assert_allclose(_sub(v.detach().cpu()).numpy(), g["s%d_w_%s" % (step, k)], rtol=0, atol=2e-5,
                                 err_msg="step %d %s" % (step, k))
  assert learner.log_losses()[1] > 0


@pytest.mark.parametrize("case", CASES)
def test_fused_gradients_match_autograd(case):
  """Every parameter gradient of one step against torch autograd through FCNetworkTrain.unroll + the same fused
  loss: float32 sums in another order (atomics), so 2e-4 of the gradient's scale per tensor."""
  from model_based_rl_b200 import fused_learner, learners
  torch.backends.cuda.matmul.allow_tf32 = False
  g = helpers.load("learner_" + case)
  cfg = _config(g)
  net, ref = _nets(g, cfg)
  fused = fused_learner.FusedLearner(cfg, net, use_graph=False, precision="f32")
  fused._stage(_batch(g, 0))
  fused._forward_and_heads_backward()
  fused._recurrent_backward()
  torch_learner = learners.Learner(cfg, ref)
  inputs = (fused.s_obs, fused.s_actions.long(), fused.s_tv, fused.s_tr, fused.s_tp, fused.s_isw)
  losses, errs = torch_learner._forward_backward(*inputs)
  torch.cuda.synchronize()
  np.testing.assert_allclose(fused.losses.cpu().numpy(), losses.cpu().numpy(), rtol=1e-5)
  np.testing.assert_allclose(fused.new_errors.cpu().numpy(), errs.cpu().numpy(), rtol=0, atol=1e-4)
  for name, p in ref.named_parameters():
    want, got = p.grad.cpu().numpy(), net.grads[name].cpu().numpy()
    scale = np.abs(want).max() + 1e-12
    assert np.abs(got - want).max() <= 2e-4 * scale, (name, np.abs(got - want).max(), scale)


def _rel(got, want):
  return float((got.double() - want.double()).abs().max() / (want.double().abs().max() + 1e-12))


def _close(got, want, bar):
  """All but one element in a thousand within `bar` of the largest entry, every element within 20 x bar: a hidden
  unit whose pre-activation lands within float32 rounding of zero is gated differently by a float32 and a float64
  sum (one such unit in ~10^6 with these shapes), which moves the few gradient entries fed by that unit."""
  err = (got.detach().double() - want.detach().double()).abs().flatten() / (want.detach().double().abs().max() + 1e-12)
  k = max(1, int(err.numel() * 0.999))
  return float(err.kthvalue(k).values) <= bar and float(err.max()) <= 20 * bar


class _EmuHead(torch.autograd.Function):
  """One head with the tensor-core kernels' rounding points (csrc/mz_learner_tc.cu) in float64: bf16 X, W1, W2, the
  512-wide activation, dY and the masked dH; exact sums (the kernels accumulate in float32); biases and their
  gradients from the unrounded values."""

  @staticmethod
  def forward(ctx, x, w1, b1, w2, b2):
    bf = lambda t: t.float().to(torch.bfloat16).double()
    xb, w1b, w2b = bf(x), bf(w1), bf(w2)
    h = torch.relu(xb @ w1b.t() + b1)
    hb = bf(h)
    ctx.save_for_backward(xb, w1b, w2b, hb, h > 0)
    return hb @ w2b.t() + b2

  @staticmethod
  def backward(ctx, dy):
    bf = lambda t: t.float().to(torch.bfloat16).double()
    xb, w1b, w2b, hb, mask = ctx.saved_tensors
    dyb = bf(dy)
    dh = (dyb @ w2b) * mask
    dhb = bf(dh)
    return dhb @ w1b, dhb.t() @ xb, dh.sum(0), dyb.t() @ hb, dy.sum(0)


def _teacher_forced_check(fused, net, fwd_bar, grad_bar):
  """Every head evaluation of the step redone in float64 autograd through _EmuHead ON THE KERNELS' OWN INPUT ROWS
  (fused.xs / s_obs), so that a bf16 rounding flip in one step does not travel into the comparison of the next:
  outputs of each evaluation within fwd_bar, every parameter gradient within grad_bar of its largest entry."""
  import torch.nn.functional as F
  K, B, A = fused.K, fused.B, net.action_space
  W = {k: v.detach().double().clone().requires_grad_(True) for k, v in net.state_dict().items()}
  head = lambda x, name, out: _EmuHead.apply(x, W[name + ".fc1.weight"], W[name + ".fc1.bias"], W["%s.%s.weight" % (name, out)],
                                             W["%s.%s.bias" % (name, out)])
  xs = fused.xs.view(K + 1, B, -1)
  # the three output heads over the stacked rows; their dX is what the chain's backward starts from
  x_all = fused.xs.double().clone().requires_grad_(True)
  v = head(x_all[:, :50], "value_head", "value")
  p = head(x_all[:, :50], "policy_head", "policy")
  r = head(x_all[:K * B], "reward_head", "reward")
  assert _close(fused.v.view(-1, v.shape[1]), v, fwd_bar) and _close(fused.p.view(-1, A), p, fwd_bar)
  assert _close(fused.r.view(-1, r.shape[1])[:K * B], r, fwd_bar)
  torch.autograd.backward([v, p, r], [fused.dv.view(-1, v.shape[1]).double(), fused.dp.view(-1, A).double(),
                                      fused.dr.view(-1, r.shape[1])[:K * B].double()])
  assert _close(fused.dxs, x_all.grad, grad_bar), _rel(fused.dxs, x_all.grad)
  up = None
  for k in range(K, -1, -1):
    x_in = (xs[k - 1] if k else fused.s_obs).double().clone().requires_grad_(True)
    y = head(x_in, "transition_head" if k else "representation_head", "out")
    h = torch.relu(F.layer_norm(y, (50,), W["LN.weight"], W["LN.bias"], 1e-5))
    assert _close(fused.yall[k], y, fwd_bar) and _close(xs[k][:, :50], h, fwd_bar), k
    grad_h = fused.dxs.view(K + 1, B, -1)[k][:, :50].double() + (up if up is not None else 0.0)
    h.backward(grad_h * (0.5 if k else 1.0))
    up = x_in.grad[:, :50]
  worst = {k: _rel(net.grads[k], W[k].grad) for k in W}
  assert all(_close(net.grads[k], W[k].grad, grad_bar) for k in W), worst
  return worst


@pytest.mark.parametrize("case", CASES)
def test_bf16_step_equals_autograd_with_the_same_rounding_points(case):
  """The four tensor-core launches of a step against torch autograd in float64 through a head that rounds to bf16
  exactly where the kernels do, evaluation by evaluation on the kernels' own inputs: outputs within 2e-3 of their
  largest entry, parameter gradients (with the kernels' own logit gradients fed back) within 3e-3 -- what is left is
  float32 accumulation, atomics, and the rare element whose float32 and float64 sums straddle a bf16 boundary."""
  from model_based_rl_b200 import fused_learner
  g = helpers.load("learner_" + case)
  cfg = _config(g)
  net, _ = _nets(g, cfg)
  fused = fused_learner.FusedLearner(cfg, net, use_graph=False)
  fused._stage(_batch(g, 0))
  fused._forward_and_heads_backward()
  fused._recurrent_backward()
  torch.cuda.synchronize()
  worst = _teacher_forced_check(fused, net, 2e-3, 3e-3)
  print(case, "vs emulation", {k: round(x, 5) for k, x in worst.items()})


@pytest.mark.parametrize("case", CASES)
def test_bf16_step_against_f32_step_and_reference_golden(case):
  """The tensor-core step against the float32 kernels on the same weights and batch: losses 1e-2 relative (and the
  same bar against the reference golden), priority errors 0.05 absolute, logits and hidden states within 3e-2 of
  their largest entry; parameter gradients as directions: cosine >= 0.98 for every tensor (8-bit mantissas through
  two to twelve chained layers, ReLU units whose pre-activation changes sign under the rounding)."""
  from model_based_rl_b200 import fused_learner
  g = helpers.load("learner_" + case)
  cfg = _config(g)
  out = {}
  for prec in ("f32", "bf16"):
    net, _ = _nets(g, cfg)
    fused = fused_learner.FusedLearner(cfg, net, use_graph=False, precision=prec)
    fused._stage(_batch(g, 0))
    fused._forward_and_heads_backward()
    fused._recurrent_backward()
    torch.cuda.synchronize()
    out[prec] = (fused, net)
  f32, bf = out["f32"][0], out["bf16"][0]
  np.testing.assert_allclose(bf.losses.cpu().numpy(), f32.losses.cpu().numpy(), rtol=1e-2)
  np.testing.assert_allclose(bf.losses.cpu().numpy(), g["s0_losses"], rtol=1e-2)
  np.testing.assert_allclose(bf.new_errors.cpu().numpy(), f32.new_errors.cpu().numpy(), rtol=0, atol=0.05)
  assert _rel(bf.xs, f32.xs) <= 3e-2 and _rel(bf.yall, f32.yall) <= 3e-2
  for name in ("v", "r", "p"):
    assert _rel(getattr(bf, name), getattr(f32, name)) <= 3e-2, name
  cos = {}
  for k, want in out["f32"][1].grads.items():
    a, b = out["bf16"][1].grads[k].double().flatten(), want.double().flatten()
    cos[k] = float((a @ b) / (a.norm() * b.norm() + 1e-300))
  print(case, "cosine vs f32", {k: round(x, 4) for k, x in cos.items()})
  assert min(cos.values()) >= 0.98, cos


@pytest.mark.parametrize("graph", [False, True], ids=["eager", "cuda_graph"])
@pytest.mark.parametrize("case", CASES)
def test_bf16_learner_two_steps_against_reference_golden(case, graph):
  """Two optimiser steps of the bf16 learner against the reference's golden run: losses 2e-2 relative; weights within
  2.5 learning rates per step taken (Adam / RMSprop normalise the update to about one learning rate per weight, so a
  weight whose tiny gradient changes sign under bf16 rounding moves by up to that much), SGD within 2e-3."""
  from model_based_rl_b200 import fused_learner
  g = helpers.load("learner_" + case)
  cfg = _config(g)
  net, _ = _nets(g, cfg)
  learner = fused_learner.FusedLearner(cfg, net, use_graph=graph)
  assert learner.precision == "bf16"
  for step in range(2):
    losses = learner.update_weights(_batch(g, step))
    np.testing.assert_allclose(losses.cpu().numpy(), g["s%d_losses" % step], rtol=2e-2)
    bar = 2e-3 if cfg.optimizer == "SGD" else 2.5 * cfg.lr_init * (step + 1)
    for k, v in net.state_dict().items():
      np.testing.assert_allclose(_sub(v.detach().cpu()).numpy(), g["s%d_w_%s" % (step, k)], rtol=0, atol=bar,
                                 err_msg="step %d %s" % (step, k))


def test_tc_head_kernels_ragged_rows_match_torch():
  """mz_heads_forward_tc / mz_heads_backward_tc on row counts that are not multiples of the 32-row tile, input widths
  that need padded k-steps, strided rows, two jobs of different length in one launch: against the same head in torch
  float64 on the bf16-rounded operands (the kernel rounds X, W, the 512-wide activation and dY / dH to bf16)."""
  import ctypes as C
  from model_based_rl_b200 import _lib
  lib = _lib.load()
  rng = torch.Generator(device="cuda").manual_seed(11)
  bf = lambda t: t.to(torch.bfloat16).double()
  jobs, keep = [], []
  for rows, d_in, d_out, ldx in ((1, 9, 9, 9), (45, 54, 50, 54), (100, 128, 31, 130), (33, 68, 4, 70)):
    X = torch.randn(rows, ldx, device="cuda", generator=rng)
    W1, b1 = torch.randn(512, d_in, device="cuda", generator=rng) * 0.1, torch.randn(512, device="cuda", generator=rng) * 0.1
    W2, b2 = torch.randn(d_out, 512, device="cuda", generator=rng) * 0.1, torch.randn(d_out, device="cuda", generator=rng)
    dY = torch.randn(rows, d_out, device="cuda", generator=rng)
    Y = torch.full((rows, d_out), 7.0, device="cuda")
    dX = torch.ones(rows, ldx, device="cuda")
    grads = [torch.zeros_like(t) for t in (W1, b1, W2, b2)]
    imgs, pj = [], []
    for (n, k), (src, sn, sk) in zip([(512, d_in), (d_out, 512), (512, d_out), (d_in, 512)],
                                     [(W1, d_in, 1), (W2, 512, 1), (W2, 1, 512), (W1, 1, d_in)]):
      img = torch.zeros(int(lib.mz_learner_packed_words(n, k)), dtype=torch.int32, device="cuda")
      imgs.append(img)
      pj.append(_lib.PackJob(src.data_ptr(), img.data_ptr(), n, k, sn, sk))
    _lib.check(lib.mz_learner_pack(4, (_lib.PackJob * 4)(*pj), None), "pack")
    head = _lib.TcHead(*[i.data_ptr() for i in imgs], b1.data_ptr(), b2.data_ptr(), *[t.data_ptr() for t in grads], d_in, d_out)
    jobs.append(_lib.TcJob(head, rows, ldx, d_out, ldx, X.data_ptr(), Y.data_ptr(), dY.data_ptr(), dX.data_ptr()))
    keep.append((rows, d_in, d_out, X, W1, b1, W2, b2, dY, Y, dX, grads, imgs))
  for first in (0, 2):  # two jobs per launch
    arr = (_lib.TcJob * 2)(*jobs[first:first + 2])
    _lib.check(lib.mz_heads_forward_tc(2, arr, None), "fwd")
    _lib.check(lib.mz_heads_backward_tc(2, arr, None), "bwd")
  torch.cuda.synchronize()
  for rows, d_in, d_out, X, W1, b1, W2, b2, dY, Y, dX, grads, _ in keep:
    x64 = bf(X[:, :d_in]).requires_grad_(True)
    w1, w2 = bf(W1).requires_grad_(True), bf(W2).requires_grad_(True)
    bb1, bb2 = b1.double().requires_grad_(True), b2.double().requires_grad_(True)
    y64 = torch.relu(x64 @ w1.t() + bb1) @ w2.t() + bb2
    y64.backward(bf(dY))
    assert _rel(Y, y64) <= 1e-2, (rows, d_in, d_out, _rel(Y, y64))
    for got, want, name in ((dX[:, :d_in] - 1.0, x64.grad, "dX"), (grads[0], w1.grad, "gW1"), (grads[1], bb1.grad, "gb1"),
                            (grads[2], w2.grad, "gW2"), (grads[3], bb2.grad, "gb2")):
      assert _rel(got, want) <= 1.5e-2, (rows, d_in, d_out, name, _rel(got, want))
    assert (dX[:, d_in:] == 1.0).all()  # columns behind d_in untouched


@pytest.mark.parametrize("B,K,A,D", [(45, 3, 4, 8), (64, 5, 18, 128), (7, 1, 9, 27)])
def test_tc_chain_kernels_match_the_per_phase_float32_kernels(B, K, A, D):
  """mz_chain_forward_tc / mz_chain_backward_tc (whole recurrent part in one launch each) against the float32
  per-phase launches on ragged batches, A = 18 (a 68-wide dynamics input: three padded k-blocks) and K = 1."""
  from model_based_rl_b200 import fused_learner
  cfg = types.SimpleNamespace(value_support=[-15, 15], reward_support=[-15, 15], no_support=False, no_target_transform=False,
                              num_unroll_steps=K, optimizer="SGD", lr_init=0.01, momentum=0.9, weight_decay=0.0, clip_grad=0,
                              lr_scheduler=None, norm_obs=False)
  r = np.random.default_rng(B)
  pol = r.random((B, K + 1, A)).astype(np.float32)
  batch = ((r.normal(size=(B, D)).astype(np.float32), r.integers(0, A, size=(B, K)).tolist(),
            ((r.random((B, K + 1)) < 0.2).astype(np.float32), (3 * r.normal(size=(B, K + 1))).astype(np.float32),
             pol / pol.sum(-1, keepdims=True))), None, r.random(B))
  out = {}
  w = None
  torch.manual_seed(1000 + B)
  for prec in ("f32", "bf16"):
    net = fused_learner.FusedFCNetwork(D, A, "cuda", cfg)
    if w is None:
      w = {k: v.clone() for k, v in net.state_dict().items()}
    net.load_weights(w)
    fused = fused_learner.FusedLearner(cfg, net, use_graph=False, precision=prec)
    fused._stage(batch)
    fused._forward_and_heads_backward()
    fused._recurrent_backward()
    torch.cuda.synchronize()
    out[prec] = (fused, net)
  f32, bf = out["f32"][0], out["bf16"][0]
  assert _rel(bf.xs, f32.xs) <= 3e-2 and _rel(bf.yall, f32.yall) <= 3e-2 and _rel(bf.rstd, f32.rstd) <= 3e-2
  assert torch.equal(bf.xs[:, 50:], f32.xs[:, 50:])  # one-hot actions
  np.testing.assert_allclose(bf.losses.cpu().numpy(), f32.losses.cpu().numpy(), rtol=1e-2)
  for k, want in out["f32"][1].grads.items():
    a, b = out["bf16"][1].grads[k].double().flatten(), want.double().flatten()
    assert float((a @ b) / (a.norm() * b.norm() + 1e-300)) >= 0.98, k
  _teacher_forced_check(bf, out["bf16"][1], 2e-3, 3e-3)


def _dp_fused_worker(rank, world_size, port, optimizer, graph, out):
  """One of two data-parallel FusedLearner ranks (both on cuda:0; gloo moves the CUDA buffers, NCCL refuses two ranks
  on one device): rank r trains on the golden batch of step r."""
  import os
  import torch.distributed as dist
  os.environ["MASTER_ADDR"] = "127.0.0.1"
  os.environ["MASTER_PORT"] = str(port)
  dist.init_process_group("gloo", rank=rank, world_size=world_size)
  try:
    from model_based_rl_b200 import fused_learner
    g = helpers.load("learner_ttt")
    cfg = _config(g)
    cfg.optimizer, cfg.clip_grad = optimizer, 0
    net, _ = _nets(g, cfg)
    if rank == 1:  # ranks start from different weights: send_weights must align them
      net.flat.add_(0.01)
    learner = fused_learner.FusedLearner(cfg, net, use_graph=graph)
    learner.send_weights()
    w0 = _weights(g)
    for k, v in net.state_dict().items():
      assert torch.equal(v.cpu(), w0[k]), k
    learner.update_weights(_batch(g, rank))
    torch.cuda.synchronize()
    mine = {k: v.detach().cpu().clone() for k, v in net.state_dict().items()}
    # single-process restatement: the gradients of both batches summed, the same optimiser step with scale 1 / 2
    solo_net, _ = _nets(g, cfg)
    solo = fused_learner.FusedLearner(cfg, solo_net, use_graph=False)
    total = torch.zeros_like(solo_net.grad)
    for r in range(world_size):
      solo._stage(_batch(g, r))
      solo._forward_and_heads_backward()
      solo._recurrent_backward()
      total += solo_net.grad
    solo_net.grad.copy_(total)
    solo._apply(world_size)
    torch.cuda.synchronize()
    bar = 1e-6 if optimizer == "SGD" else 2.5 * cfg.lr_init  # atomics order the sums differently; Adam normalises
    for k, v in solo_net.state_dict().items():
      assert torch.allclose(v.cpu(), mine[k], rtol=0, atol=bar), (k, float((v.cpu() - mine[k]).abs().max()))
    gathered = [None] * world_size
    dist.all_gather_object(gathered, {k: v.numpy() for k, v in mine.items()})
    for k in mine:
      assert np.array_equal(gathered[0][k], gathered[1][k]), k  # every rank holds the same weights, bit for bit
    out.put((rank, "ok"))
  except Exception:  # pragma: no cover
    import traceback
    out.put((rank, traceback.format_exc()))
  finally:
    dist.destroy_process_group()


@pytest.mark.parametrize("optimizer,graph", [("SGD", False), ("AdamW", True)])
def test_fused_learner_data_parallel_two_ranks(optimizer, graph):
  """The bucketed gradient all-reduce of FusedLearner (output heads' bucket on a side stream under the recurrent
  backward, the rest after it) with two ranks: same weights on both ranks bit for bit, equal to one process applying
  the summed gradients of both batches."""
  import socket
  import torch.multiprocessing as mp
  ctx = mp.get_context("spawn")
  out = ctx.Queue()
  sk = socket.socket()
  sk.bind(("127.0.0.1", 0))
  port = sk.getsockname()[1]
  sk.close()
  procs = [ctx.Process(target=_dp_fused_worker, args=(r, 2, port, optimizer, graph, out)) for r in range(2)]
  for p in procs:
    p.start()
  res = [out.get(timeout=300) for _ in procs]
  for p in procs:
    p.join(timeout=60)
  assert sorted(res) == [(0, "ok"), (1, "ok")], res


def test_mlp2_kernels_ragged_rows_match_torch():
  """mz_mlp2_forward / mz_mlp2_backward on row counts that are not multiples of the 32-row tile, wide inputs, strided
  rows: against the same two-layer head in torch (float64 reference)."""
  import ctypes as C
  from model_based_rl_b200 import _lib
  lib = _lib.load()
  P = lambda t: C.c_void_p(t.data_ptr())
  rng = torch.Generator(device="cuda").manual_seed(5)
  for rows, d_in, d_out, ldx in ((1, 9, 9, 9), (45, 54, 50, 54), (100, 128, 31, 130), (33, 50, 4, 59)):
    X = torch.randn(rows, ldx, device="cuda", generator=rng)
    W1, b1 = torch.randn(512, d_in, device="cuda", generator=rng) * 0.1, torch.randn(512, device="cuda", generator=rng) * 0.1
    W2, b2 = torch.randn(d_out, 512, device="cuda", generator=rng) * 0.1, torch.randn(d_out, device="cuda", generator=rng)
    W1T, W2T = W1.t().contiguous(), W2.t().contiguous()
    Y = torch.full((rows, d_out), 7.0, device="cuda")
    _lib.check(lib.mz_mlp2_forward(rows, d_in, ldx, P(X), P(W1T), P(b1), P(W2T), P(b2), d_out, P(Y), d_out, None), "fwd")
    x64 = X[:, :d_in].double().requires_grad_(True)
    w1, bb1, w2, bb2 = [t.double().requires_grad_(True) for t in (W1, b1, W2, b2)]
    y64 = torch.relu(x64 @ w1.t() + bb1) @ w2.t() + bb2
    torch.cuda.synchronize()
    assert torch.allclose(Y.double(), y64, rtol=1e-4, atol=1e-4), (rows, d_in, d_out)
    dY = torch.randn(rows, d_out, device="cuda", generator=rng)
    y64.backward(dY.double())
    dX = torch.ones(rows, ldx, device="cuda")  # accumulated into
    gW1, gb1, gW2, gb2 = torch.zeros_like(W1), torch.zeros_like(b1), torch.zeros_like(W2), torch.zeros_like(b2)
    _lib.check(lib.mz_mlp2_backward(rows, d_in, ldx, P(X), P(W1T), P(b1), P(W1), P(W2), d_out, P(dY), d_out, P(dX), ldx,
                                    P(gW1), P(gb1), P(gW2), P(gb2), None), "bwd")
    torch.cuda.synchronize()
    for got, want, name in ((dX[:, :d_in] - 1.0, x64.grad, "dX"), (gW1, w1.grad, "gW1"), (gb1, bb1.grad, "gb1"),
                            (gW2, w2.grad, "gW2"), (gb2, bb2.grad, "gb2")):
      scale = want.abs().max().item() + 1e-12
      assert (got.double() - want).abs().max().item() <= 1e-4 * scale, (rows, d_in, d_out, name)
    assert (dX[:, d_in:] == 1.0).all()  # columns behind d_in untouched


@pytest.mark.parametrize("precision", ["bf16", "f32"])
def test_fused_learner_with_replay_feedback_and_search_hand_off(precision):
  """learn(): sample_batch_device(ring) -> FusedLearner.update_weights -> priorities fed back on the device, and
  the weight hand-off into the search kernels' FCNetwork."""
  from model_based_rl_b200 import fused_learner
  from model_based_rl_b200.networks import FCNetwork
  from model_based_rl_b200.replay_buffer import PrioritizedReplay
  from model_based_rl_b200.selfplay import HistorySlice
  D, A, K = 8, 4, 3
  cfg = types.SimpleNamespace(value_support=[-15, 15], reward_support=[-15, 15], no_support=False, no_target_transform=False,
                              num_unroll_steps=K, td_steps=5, optimizer="AdamW", lr_init=0.001, momentum=0.9,
                              weight_decay=1e-4, clip_grad=0, lr_scheduler="ExponentialLR", lr_decay_rate=0.99, norm_obs=False,
                              batch_size=64, beta_increment_per_sampling=0.001, epsilon=0.01, alpha=1.0, beta=0.5, discount=0.997,
                              action_space=A, obs_space=(D,), window_size=2000, window_step=None, seed=None,
                              send_weights_frequency=3, training_steps=7, stored_before_train=100)
  rb = PrioritizedReplay(cfg)
  r2 = np.random.default_rng(5)
  for _ in range(8):
    n = int(r2.integers(50, 300))
    rb.save_history(HistorySlice(list(r2.normal(size=(n, D)).astype(np.float32)), r2.random((n, A)).tolist(),
                                 r2.normal(size=n).tolist(), r2.integers(0, A, size=n).tolist(), r2.normal(size=n).tolist(),
                                 np.abs(r2.normal(size=n)).tolist(), [False] * n, list(range(n)), [None] * n, [1] * n),
                    ignore=None, terminal=True)
  net = fused_learner.FusedFCNetwork(D, A, "cuda", cfg)
  search_net = FCNetwork(D, A, "cuda", cfg)
  learner = fused_learner.FusedLearner(cfg, net, replay_buffer=rb, search_network=search_net, precision=precision)
  before = rb.index.tree.clone()
  w0 = net.flat.clone()
  assert learner.learn() == 7
  torch.cuda.synchronize()
  assert not torch.equal(rb.index.tree, before)          # priorities came back
  assert not torch.equal(net.flat, w0) and torch.isfinite(net.flat).all()
  assert abs(learner.lr_scheduler.lr - 0.001 * 0.99 ** 7) < 1e-12
  got = search_net.get_weights()
  learner.send_weights()
  for k, v in net.state_dict().items():
    assert torch.equal(search_net.get_weights()[k].cpu(), v.cpu()), k
  assert set(got) == set(net.state_dict())
