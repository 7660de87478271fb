"""GPU parity tests of the tensor-core MuZeroNetwork path (csrc/mz_conv_tc.cu) against the float32
oracle (oracle/muzero_ref.py, pinned to the reference class by tests/test_oracle_golden.py) and the
reference golden outputs.  Tolerances: bf16 operands (8-bit mantissa) through 33 / 65 layers with
float32 accumulation -- stated per assertion."""
import types

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from helpers import load

pytestmark = pytest.mark.gpu

CFG = types.SimpleNamespace(value_support=[-15, 15], reward_support=[-15, 15], no_support=False,
                            no_target_transform=False)


def _bf16(x):
  return x.to(torch.bfloat16).float()


@pytest.mark.parametrize("games,flags", [(1, 1), (2, 1), (7, 3), (300, 3), (5, 1 | 4), (4, 1 | 2 | 8)])
def test_conv_kernel_matches_conv2d(games, flags):
  """One launch of the implicit GEMM vs F.conv2d on the same bf16-rounded operands (float32 math):
  only the accumulation order differs -> 2e-2 absolute on O(1) outputs after bf16 rounding of the
  result.  Game counts that do not fill a 128-row tile, tiles that straddle games."""
  from model_based_rl_b200 import _lib, muzero
  lib = _lib.load()
  torch.manual_seed(games * 8 + flags)
  dev = "cuda"
  x = torch.rand((games, 128, 6, 6), device=dev)
  cin = 129 if flags & 4 else 128
  w = (torch.rand((128, cin, 3, 3), device=dev) * 2 - 1) * (3.0 / (cin * 9)) ** 0.5
  bias = torch.randn(128, device=dev) * 0.1
  conv = muzero._Conv(w, bias, None, dev)
  rows = muzero.to_padded(x)
  res = torch.rand((games, 128, 6, 6), device=dev)
  res_rows = muzero.to_padded(res)
  actions = torch.randint(0, 18, (games,), device=dev, dtype=torch.int32)
  out = torch.full((games * 49, 128), 7.0, dtype=torch.bfloat16, device=dev)
  scaled = torch.full((games * 49, 128), 7.0, dtype=torch.bfloat16, device=dev)
  P = _lib.ptr
  _lib.check(lib.mz_conv3x3_tc(games, 6, 128, P(rows), P(conv.w), P(conv.bias), flags,
                               P(conv.plane) if flags & 4 else None, P(actions), 18,
                               P(res_rows) if flags & 2 else None, P(out),
                               P(scaled) if flags & 8 else None, None, None, _lib.current_stream()), "conv")
  torch.cuda.synchronize()
  xin = _bf16(x)
  wq = _bf16(conv.w.float().reshape(128, 3, 3, 128).permute(0, 3, 1, 2))
  want = F.conv2d(xin, wq, None, 1, 1) + conv.bias[None, :, None, None]
  if flags & 4:
    plane = conv.plane.reshape(6, 6, 128).permute(2, 0, 1)
    want = want + (actions.float() / 18)[:, None, None, None] * plane[None]
  if flags & 2:
    want = want + _bf16(res)
  if flags & 1:
    want = F.relu(want)
  got = muzero.from_padded(out, games)
  assert torch.allclose(got, want, rtol=1e-2, atol=2e-2), float((got - want).abs().max())
  assert float(muzero.padding_rows(out, games).float().abs().max()) == 0  # the padding stays zero
  if flags & 8:
    mn, mx = want.min(dim=1, keepdim=True)[0], want.max(dim=1, keepdim=True)[0]
    want_s = (want - mn) / (mx - mn)
    got_s = muzero.from_padded(scaled, games)
    assert torch.allclose(got_s, want_s, rtol=1e-2, atol=2e-2), float((got_s - want_s).abs().max())
    assert float(muzero.padding_rows(scaled, games).float().abs().max()) == 0


@pytest.mark.parametrize("games,width,flags", [(1, 6, 1), (2, 6, 3), (3, 6, 3), (301, 6, 1 | 2 | 8), (5, 6, 1 | 4),
                                               (3, 12, 3), (2, 24, 1)])
def test_conv_pair_kernel_equals_single_cta_kernel(games, width, flags):
  """The CTA-pair kernel (tcgen05.mma.cta_group::2, M = 256, resident weights), with one activation load
  per tap (mode 1) and with one row window per tile (mode 2), against the single-CTA kernel on the same operands: same K order and float32 accumulation, same epilogue -> identical bits.
  Odd tile counts (the pair's second half past the end), tiles straddling games, every epilogue."""
  from model_based_rl_b200 import _lib, muzero
  lib = _lib.load()
  torch.manual_seed(games * 31 + flags + width)
  dev = "cuda"
  wp = width + 1
  x = torch.rand((games, 128, width, width), device=dev)
  cin = 129 if flags & 4 else 128
  w = (torch.rand((128, cin, 3, 3), device=dev) * 2 - 1) * (3.0 / (cin * 9)) ** 0.5
  conv = muzero._Conv(w, torch.randn(128, device=dev) * 0.1, None, dev)
  rows = muzero.to_padded(x)
  res_rows = muzero.to_padded(torch.rand((games, 128, width, width), device=dev))
  actions = torch.randint(0, 18, (games,), device=dev, dtype=torch.int32)
  P = _lib.ptr
  outs = []
  try:
    for pair in (0, 1, 2):
      assert lib.mz_conv_set_pair(pair) == 0
      out = torch.full((games * wp * wp, 128), 7.0, dtype=torch.bfloat16, device=dev)
      scaled = torch.full((games * wp * wp, 128), 7.0, dtype=torch.bfloat16, device=dev)
      for _ in range(2):  # the second launch reuses barriers / TMEM of a warm SM
        _lib.check(lib.mz_conv3x3_tc(games, width, 128, P(rows), P(conv.w), P(conv.bias), flags,
                                     P(conv.plane) if flags & 4 else None, P(actions), 18,
                                     P(res_rows) if flags & 2 else None, P(out),
                                     P(scaled) if flags & 8 else None, None, None, _lib.current_stream()), "conv")
      torch.cuda.synchronize()
      outs.append((out, scaled))
  finally:
    lib.mz_conv_set_pair(2)
  for mode in (1, 2):
    assert torch.equal(outs[0][0], outs[mode][0]), (mode, float((outs[0][0].float() - outs[mode][0].float()).abs().max()))
    if flags & 8:
      assert torch.equal(outs[0][1], outs[mode][1]), mode


@pytest.mark.parametrize("width,games,ch", [(12, 3, 128), (24, 2, 128), (48, 2, 64)])
def test_conv_kernel_other_widths(width, games, ch):
  """The same kernel on the 12 x 12 / 24 x 24 (128 channels) and 48 x 48 (64 channels) stages of the
  representation tower."""
  from model_based_rl_b200 import _lib, muzero
  lib = _lib.load()
  torch.manual_seed(width)
  dev = "cuda"
  x = torch.rand((games, ch, width, width), device=dev)
  w = (torch.rand((ch, ch, 3, 3), device=dev) * 2 - 1) * (3.0 / (9 * ch)) ** 0.5
  conv = muzero._Conv(w, torch.randn(ch, device=dev) * 0.1, None, dev)
  rows = muzero.to_padded(x)
  out = torch.full_like(rows, 7.0)
  P = _lib.ptr
  _lib.check(lib.mz_conv3x3_tc(games, width, ch, P(rows), P(conv.w), P(conv.bias), 1 | 2, None, None, 18, P(rows),
                               P(out), None, None, None, _lib.current_stream()), "conv")
  torch.cuda.synchronize()
  want = F.relu(F.conv2d(_bf16(x), _bf16(w), None, 1, 1) + conv.bias[None, :, None, None] + _bf16(x))
  got = muzero.from_padded(out, games, width)
  assert torch.allclose(got, want, rtol=1e-2, atol=2e-2), float((got - want).abs().max())
  assert float(muzero.padding_rows(out, games, width).float().abs().max()) == 0


@pytest.mark.parametrize("cin,cout,w_in,games", [(4, 64, 96, 2), (32, 64, 96, 1), (64, 128, 48, 3)])
def test_strided_conv_and_avgpool(cin, cout, w_in, games):
  """Conv2d(stride 2) as im2col + GEMM into the padded layout, and AvgPool2d(3, 2, 1), vs torch on the
  same bf16-rounded operands."""
  from model_based_rl_b200 import muzero
  dev = "cuda"
  torch.manual_seed(cin + w_in)
  net = muzero.MuZeroNetwork(cin, 18, dev, CFG)
  x = torch.rand((games, cin, w_in, w_in), device=dev)
  w = (torch.rand((cout, cin, 3, 3), device=dev) * 2 - 1) * (3.0 / (9 * cin)) ** 0.5
  bias = torch.randn(cout, device=dev) * 0.1
  conv = muzero._StridedConv(w, bias, dev)
  xp = F.pad(x, (0, 0, 0, 0, 0, conv.cin_pad - cin))
  rows = muzero.to_padded(xp)
  w_out = w_in // 2
  out = torch.zeros((games * (w_out + 1) ** 2, cout), dtype=torch.bfloat16, device=dev)
  net._strided(games, conv, rows, w_in, out)
  torch.cuda.synchronize()
  want = F.conv2d(_bf16(x), _bf16(w), bias, 2, 1)
  got = muzero.from_padded(out, games, w_out)
  assert torch.allclose(got, want, rtol=1e-2, atol=2e-2), float((got - want).abs().max())
  assert float(muzero.padding_rows(out, games, w_out).float().abs().max()) == 0
  if cout == 128:
    pooled = torch.zeros((games * (w_out // 2 + 1) ** 2, 128), dtype=torch.bfloat16, device=dev)
    net._pool(games, out, w_out, pooled)
    torch.cuda.synchronize()
    want_p = F.avg_pool2d(muzero.from_padded(out, games, w_out), 3, 2, 1)
    got_p = muzero.from_padded(pooled, games, w_out // 2)
    assert torch.allclose(got_p, want_p, rtol=1e-2, atol=1e-2)
    assert float(muzero.padding_rows(pooled, games, w_out // 2).float().abs().max()) == 0


def test_pool_gather_scatter_and_fc_heads():
  """mz_conv_gather (hidden-pool gather), the pool scatter of the scaling epilogue, the Linear heads."""
  from model_based_rl_b200 import _lib, muzero
  lib = _lib.load()
  dev = "cuda"
  torch.manual_seed(5)
  games, slots = 6, 5
  pool = torch.zeros((games * slots * 49, 128), dtype=torch.bfloat16, device=dev)
  x = torch.rand((games, 128, 6, 6), device=dev)
  pick = torch.randint(0, slots, (games,), device=dev, dtype=torch.int32)
  rows = muzero.to_padded(x).reshape(games, 49, 128)
  pv = pool.view(games, slots, 49, 128)
  for g in range(games):
    pv[g, int(pick[g])] = rows[g]
  flat = torch.zeros((games * 49, 128), dtype=torch.bfloat16, device=dev)
  P = _lib.ptr
  _lib.check(lib.mz_conv_gather(games, slots, P(pick), P(pool), P(flat), _lib.current_stream()), "gather")
  torch.cuda.synchronize()
  assert torch.equal(flat, rows.reshape(-1, 128))
  w = (torch.rand((128, 128, 3, 3), device=dev) * 2 - 1) * 0.05
  conv = muzero._Conv(w, None, None, dev)
  out = torch.zeros((games * 49, 128), dtype=torch.bfloat16, device=dev)
  scaled = torch.zeros_like(out)
  dst = (pick + 1) % slots
  out_base = ((torch.arange(games, device=dev) * slots + dst) * 49).to(torch.int32)
  _lib.check(lib.mz_conv3x3_tc(games, 6, 128, P(flat), P(conv.w), P(conv.bias), 1 | 2 | 8, None, None, 18, P(flat),
                               P(out), P(scaled), P(pool), P(out_base), _lib.current_stream()), "conv")
  torch.cuda.synchronize()
  want = F.relu(F.conv2d(_bf16(x), _bf16(w), None, 1, 1) + _bf16(x))
  assert torch.allclose(muzero.from_padded(out, games), want, rtol=1e-2, atol=2e-2)
  mn, mx = want.min(dim=1, keepdim=True)[0], want.max(dim=1, keepdim=True)[0]
  assert torch.allclose(muzero.from_padded(scaled, games), (want - mn) / (mx - mn), rtol=1e-2, atol=2e-2)
  got_pool = torch.stack([pv[g, int(dst[g])] for g in range(games)]).reshape(-1, 128)
  assert torch.equal(got_pool, scaled)
  # heads: Linear(4608 -> 256) + ReLU over the padded rows, then Linear(512 -> 31) -> scalar
  fcw = (torch.rand((256, 4608), device=dev) * 2 - 1) * 0.02
  fcb = torch.randn(256, device=dev) * 0.1
  hid = torch.zeros((games, 512), dtype=torch.float32, device=dev)
  _lib.check(lib.mz_conv_fc_tc(games, P(out), P(muzero._pack_fc(fcw, dev)), P(fcb), 256, 1, P(hid), 512,
                               _lib.current_stream()), "fc")
  torch.cuda.synchronize()
  flat_in = _bf16(muzero.from_padded(out, games)).reshape(games, -1)
  want_h = F.relu(F.linear(flat_in, _bf16(fcw), fcb))
  assert torch.allclose(hid[:, :256], want_h, rtol=1e-2, atol=1e-2)
  w2 = torch.randn((31, 512), device=dev) * 0.05
  b2 = torch.randn(31, device=dev) * 0.1
  hid = torch.rand((games, 512), device=dev)
  sc = torch.zeros(games, device=dev)
  _lib.check(lib.mz_conv_head(games, P(hid), 512, P(w2), P(b2), 31, 1, -15, 0, P(sc), 1,
                              _lib.current_stream()), "head")
  torch.cuda.synchronize()
  from oracle import muzero_ref
  want_sc = muzero_ref.inverse_transform(F.linear(hid, w2, b2), -15, 15).flatten()
  assert torch.allclose(sc, want_sc, rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("fixture", ["muzero_net", "muzero_net_c32", "muzero_net_c64"])
def test_network_matches_reference_golden(fixture):
  """Whole network against outputs of the reference's own MuZeroNetwork class (eval mode): initial_inference +
  two recurrent_inference steps.  muzero_net: C_in = 4, B = 2; muzero_net_c32 / _c64: the benchmarked towers
  (stack_obs = 32, and 64 planes with stack_actions, utils.py:28-32) at B = 8, observations redrawn from the
  fixture's seeded generator."""
  from model_based_rl_b200.muzero import MuZeroNetwork
  from oracle import muzero_ref
  g = load(fixture)
  C_in, A = int(g["input_channels"]), int(g["action_space"])
  net = MuZeroNetwork(C_in, A, "cuda", CFG)
  net.load_weights(muzero_ref.seeded_state_dict(C_in, A, int(g["seed"])))
  if "obs" in g:
    obs = torch.from_numpy(g["obs"])
  else:
    gen = torch.Generator().manual_seed(int(g["seed"]) + 1)
    obs = torch.rand((int(g["batch"]), C_in, 96, 96), generator=gen)
    assert float(obs.double().sum().item()) == float(g["obs_checksum"])
  init = net.initial_inference(obs)
  # representation: float32 stem, then 22 residual blocks in bf16 on the tensor cores
  err = np.abs(init.hidden_state.cpu().numpy() - g["init_hidden"])
  print("init_hidden", float(err.max()), float(err.mean()))
  assert err.max() < 0.06 and err.mean() < 0.01
  rec = net.recurrent_inference(torch.from_numpy(g["init_hidden"]).cuda(), g["actions"].tolist())
  rec2 = net.recurrent_inference(torch.from_numpy(g["rec_hidden"]).cuda(), g["actions2"].tolist())
  torch.cuda.synchronize()
  report = {}
  for name, got, want in (("init_logits", init.policy_logits, g["init_logits"]),
                          ("init_value", init.value, g["init_value"]),
                          ("rec_hidden", rec.hidden_state, g["rec_hidden"]),
                          ("rec_logits", rec.policy_logits, g["rec_logits"]),
                          ("rec_value", rec.value, g["rec_value"]), ("rec_reward", rec.reward, g["rec_reward"]),
                          ("rec2_hidden", rec2.hidden_state, g["rec2_hidden"]),
                          ("rec2_logits", rec2.policy_logits, g["rec2_logits"]),
                          ("rec2_value", rec2.value, g["rec2_value"]),
                          ("rec2_reward", rec2.reward, g["rec2_reward"])):
    err = np.abs(got.cpu().numpy() - want)
    report[name] = (float(err.max()), float(err.mean()))
  print(report)
  # bf16 activations through 33 (dynamics) / 32 (prediction) convolutions: scaled states are in
  # [0, 1] -> 0.06 absolute worst pixel, 0.01 mean; logits +-3 -> 0.15; scalars (|v| ~ 13) -> 0.3
  assert report["rec_hidden"][0] < 0.06 and report["rec_hidden"][1] < 0.01
  assert report["rec2_hidden"][0] < 0.06 and report["rec2_hidden"][1] < 0.01
  for k in ("init_logits", "rec_logits", "rec2_logits"):
    assert report[k][0] < 0.15, (k, report[k])
  # scalars: the network's error lives in support space (softmax expectation over the 31 bins: logit errors of
  # ~0.03 move it by up to ~0.05) and h^-1 stretches it by dh^-1/dx ~ 2 sqrt(|v| + 1) (7.7 at |v| = 14), so the bar
  # is stated where it is uniform: |h(got) - h(want)| < 0.1 (mean < 0.05) with h = Config.scalar_transform
  # (config.py:51-54); measured worst 0.071 (C_in = 64, reward after two steps)
  h = lambda v: np.sign(v) * (np.sqrt(np.abs(v) + 1) - 1) + 0.001 * v
  for name, got, want in (("init_value", init.value, g["init_value"]), ("rec_value", rec.value, g["rec_value"]),
                          ("rec_reward", rec.reward, g["rec_reward"]), ("rec2_value", rec2.value, g["rec2_value"]),
                          ("rec2_reward", rec2.reward, g["rec2_reward"])):
    dh = np.abs(h(got.cpu().numpy().astype(np.float64)) - h(want.astype(np.float64)))
    assert dh.max() < 0.1 and dh.mean() < 0.05, (name, float(dh.max()), float(dh.mean()), report[name])
    assert report[name][0] < 0.45, (name, report[name])  # and never more than 0.45 in value space (|v| <= 15)


def test_conv_search_replays_bit_exact_in_oracle():
  """A full move with MuZeroNetwork in the loop (hidden pool gathered / scattered in place): the
  engine records what the network returned for every simulation; the oracle replays the search with
  those outputs and must reproduce parents, actions, visit counts and root values bit for bit.  The
  recorded outputs themselves are checked against the float32 oracle network along one game's path."""
  import oracle
  from oracle import muzero_ref
  from model_based_rl_b200.muzero import ConvSearch, MuZeroNetwork, from_padded
  C_in, A, G, S = 4, 18, 7, 12
  sd = muzero_ref.seeded_state_dict(C_in, A, 99)
  net = MuZeroNetwork(C_in, A, "cuda", CFG)
  net.load_weights(sd)
  cfg = types.SimpleNamespace(num_simulations=S, action_space=A, two_players=False, discount=0.997,
                              pb_c_base=19652, pb_c_init=1.25, init_value_score=0.0,
                              known_bounds=[None, None], root_exploration_fraction=0.25)
  rng = np.random.default_rng(3)
  obs = torch.from_numpy(rng.random((G, C_in, 96, 96)).astype(np.float32))
  noise = rng.dirichlet([0.25] * A, size=G)
  u, temp = rng.random(G), rng.choice([0.0, 1.0], size=G)
  for use_graph in (False, True):
    cs = ConvSearch(cfg, net, G, use_graph=use_graph)
    cs.enable_record()
    actions, root_value, child_visits, init_value = cs.search(obs, noise, u, temp)
    torch.cuda.synchronize()
    if use_graph:
      a2 = actions.clone()
      actions, root_value, child_visits, init_value = cs.search(obs, noise, u, temp)
      torch.cuda.synchronize()
      assert torch.equal(a2, actions)
    ocfg = oracle.make_cfg(S, A, False, 0.997)
    want = oracle.search(ocfg, cs.root_logits.cpu().numpy(), noise=noise, noise_frac=0.25,
                         rec_value=cs.record[0].cpu().numpy().T, rec_reward=cs.record[1].cpu().numpy().T,
                         rec_logits=cs.record[2].cpu().numpy().transpose(1, 0, 2))
    assert np.array_equal(cs.trace[0].cpu().numpy().T, want["trace_parent"])
    assert np.array_equal(cs.trace[1].cpu().numpy().T, want["trace_action"])
    assert np.array_equal(cs.eng.visits.cpu().numpy(), want["visits"])
    assert np.array_equal(root_value.cpu().numpy(), want["root_value"])
    for i in range(G):
      assert int(actions[i]) == oracle.select_action(want["visits"][i], temp[i], u[i])
  # the network outputs along game 0's simulations vs the float32 oracle network fed the same
  # (bf16) parent states from the pool: one recurrent step each.  value / reward are h^-1 of the
  # support expectation, whose slope grows with |x| (about 6 at |v| = 12): 0.05 + 5 % of |v|;
  # logits 0.1, scaled states 0.05
  sdc = {k: v.cuda() for k, v in sd.items()}
  pool = cs.pool.view(G, S + 1, 49, 128)
  parents, acts = cs.trace[0].cpu().numpy()[:, 0], cs.trace[1].cpu().numpy()[:, 0]
  for sim in range(S):
    h = from_padded(pool[0, int(parents[sim])].reshape(49, 128), 1)
    v, r, pol, h2 = muzero_ref.recurrent_inference(h, [int(acts[sim])], sdc, A)
    assert abs(float(v) - float(cs.record[0][sim, 0])) < 0.05 + 0.05 * abs(float(v))
    assert abs(float(r) - float(cs.record[1][sim, 0])) < 0.05 + 0.05 * abs(float(r))
    assert float((pol[0] - cs.record[2][sim, 0]).abs().max()) < 0.1
    got_h = from_padded(pool[0, sim + 1].reshape(49, 128), 1)
    assert float((got_h - h2).abs().max()) < 0.05


def test_selfplay_driver_through_conv_search():
  """BatchedActor(search=ConvSearch): the per-move body of Actor.play_game (actors.py:125-176) with
  MuZeroNetwork behind the same `search_host` call as FCSearch.  Moves are compared with a second actor
  that drives the generic path (BatchedMCTS + one recurrent_inference call per simulation) on an identical
  environment and identical draws; history slices reach the sink with the reference's chunking rules."""
  from model_based_rl_b200.environments import SyntheticFrames
  from model_based_rl_b200.muzero import ConvSearch, MuZeroNetwork
  from model_based_rl_b200.selfplay import BatchedActor
  from oracle import muzero_ref
  C_in, A, G, S = 4, 6, 5, 6
  net = MuZeroNetwork(C_in, A, "cuda", CFG)
  net.load_weights(muzero_ref.seeded_state_dict(C_in, A, 7))
  cfg = types.SimpleNamespace(num_simulations=S, action_space=A, two_players=False, discount=0.997,
                              pb_c_base=19652, pb_c_init=1.25, init_value_score=0.0, known_bounds=[None, None],
                              root_exploration_fraction=0.25, root_dirichlet_alpha=0.25, num_unroll_steps=2,
                              td_steps=2, max_history_length=3, max_steps=100, norm_obs=False)
  fast = BatchedActor(cfg, net, SyntheticFrames(G, A, (C_in, 96, 96), episode_length=4, seed=11),
                      search=ConvSearch(cfg, net, G))
  slow = BatchedActor(cfg, net, SyntheticFrames(G, A, (C_in, 96, 96), episode_length=4, seed=11))
  rng = np.random.default_rng(5)
  for move in range(6):
    noise, u = rng.dirichlet([0.25] * A, size=G), rng.random(G)
    a1, v1, cv1, e1, d1 = fast.play_move(noise=noise, uniforms=u)
    a2, v2, cv2, e2, d2 = slow.play_move(noise=noise, uniforms=u)
    # both paths evaluate the same bf16 network; the generic path round-trips hidden states through the
    # reference's [G, 128, 6, 6] float32 layout, so values agree to rounding and visit counts almost always
    assert np.allclose(v1, v2, atol=5e-2), (move, v1, v2)
    assert np.abs(cv1 - cv2).max() <= 2.0 / S + 1e-12
    assert np.array_equal(d1, d2)
    assert cv1.shape == (G, A) and np.allclose(cv1.sum(1), 1.0)
    assert ((a1 >= 0) & (a1 < A)).all()
    # keep the two environments in lock step whatever the sampled actions were
    slow.env.rng = np.random.default_rng(100 + move)
    fast.env.rng = np.random.default_rng(100 + move)
  assert fast.games_played == G and fast.experiences_collected == 6 * G
  # chunking: max_history_length = 3 -> one slice after 3 moves (ignore = overlap), one at the terminal move
  kinds = [(ign, term) for (_, _, ign, term) in fast.saved]
  assert kinds.count((4, False)) == G and kinds.count((None, True)) == G
  first = [h for (i, h, ign, term) in fast.saved if i == 0][0]
  assert len(first.actions) == 3 and len(first.observations) == 4 and first.observations[0].shape == (C_in, 96, 96)


def test_conv_search_after_load_weights_uses_new_weights():
  """MuZeroNetwork.load_weights rebuilds the packed tensors; a ConvSearch whose CUDA graph was captured before
  the hand-off must re-capture, not replay launches that read the freed allocations: the search after the
  load equals the search of an engine built afterwards."""
  from model_based_rl_b200.muzero import ConvSearch, MuZeroNetwork, random_state_dict
  G, A, S, C_in = 8, 6, 4, 4
  cfg = types.SimpleNamespace(num_simulations=S, action_space=A, two_players=False, discount=0.997, pb_c_base=19652,
                              pb_c_init=1.25, init_value_score=0.0, known_bounds=[None, None],
                              root_exploration_fraction=0.25, value_support=[-15, 15], reward_support=[-15, 15],
                              no_support=False, no_target_transform=False)
  rng = np.random.default_rng(4)
  obs = torch.from_numpy(rng.random((G, C_in, 96, 96), dtype=np.float32))
  noise, u, temp = rng.dirichlet([0.25] * A, size=G), rng.random(G), np.ones(G)
  net = MuZeroNetwork(C_in, A, "cuda", cfg)
  net.load_weights(random_state_dict(C_in, A, seed=1))
  cs = ConvSearch(cfg, net, G, use_graph=True)
  first = [t.clone() for t in cs.search(obs, noise, u, temp)]
  sd2 = random_state_dict(C_in, A, seed=2)
  net.load_weights(sd2)
  second = [t.clone() for t in cs.search(obs, noise, u, temp)]
  fresh_net = MuZeroNetwork(C_in, A, "cuda", cfg)
  fresh_net.load_weights(sd2)
  want = [t.clone() for t in ConvSearch(cfg, fresh_net, G, use_graph=False).search(obs, noise, u, temp)]
  torch.cuda.synchronize()
  assert not torch.equal(first[3], second[3])
  for x, y in zip(second, want):
    assert torch.equal(x, y)
