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
  learner = fused_learner.FusedLearner(cfg, net, use_graph=graph)
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
  fused = fused_learner.FusedLearner(cfg, net, use_graph=False)
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


def test_fused_learner_with_replay_feedback_and_search_hand_off():
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
  learner = fused_learner.FusedLearner(cfg, net, replay_buffer=rb, search_network=search_net)
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
