"""Pins the CPU oracle (oracle/mz_oracle.c) to outputs of the unmodified reference (tests/golden).

Everything the search produces is compared BIT FOR BIT, including float64 priors, value sums and
MinMaxStats: on this platform the oracle's libm exp/log are the ones math.exp/math.log call.
"""
import numpy as np
import pytest

import oracle
from helpers import (REPLAY_CASES, SEARCH_CASES, SEARCH_EXACT_KEYS, is_clip_case, load,
                     oracle_search_from_golden)


@pytest.mark.parametrize("case", SEARCH_CASES)
def test_search_bit_exact(case):
  g = load("search_" + case)
  out = oracle_search_from_golden(g)
  for k in SEARCH_EXACT_KEYS:
    want, got = g[k], out[k]
    if k == "edge_child":
      pass
    assert want.shape == got.shape, k
    assert np.array_equal(want, got), "%s differs (max abs %g)" % (
        k, np.max(np.abs(want.astype(np.float64) - got.astype(np.float64))))


def test_py_sum_matches_builtin():
  rng = np.random.default_rng(1)
  for _ in range(3000):
    n = int(rng.integers(1, 19))
    xs = np.exp(rng.normal(0, 2, size=n).astype(np.float32).astype(np.float64))
    assert oracle.py_sum(xs) == sum(float(x) for x in xs)
    s = 0.0
    for x in xs:
      s += float(x)
    assert oracle.py_sum(xs, mode=0) == s


def test_select_action_and_child_visits():
  g = load("select_action")
  for i in range(len(g["A"])):
    A, mask = int(g["A"][i]), int(g["legal"][i])
    actions = [a for a in range(A) if (mask >> a) & 1]
    dense = g["visits"][i][actions]
    idx = oracle.select_action(dense, g["temperature"][i], g["u"][i])
    assert actions[idx] == int(g["action"][i]), i
    child = np.array([-1 if (mask >> a) & 1 else -2 for a in range(A)], np.int32)
    cv = oracle.child_visits(g["visits"][i][:A], child)
    assert np.array_equal(cv, g["child_visits"][i][:A]), i
    assert g["root_vsum"][i] / g["root_visit"][i] == g["root_value"][i]


@pytest.mark.parametrize("case", REPLAY_CASES)
def test_insert_target(case):
  g = load("replay_" + case)
  K, T, disc = int(g["num_unroll_steps"]), int(g["td_steps"]), float(g["discount"])
  worst = 0.0
  for b in range(int(g["n_batches"])):
    steps, hist = g["b%d_steps" % b], g["b%d_hist" % b]
    for row in range(len(steps)):
      h = int(hist[row])
      rewards = g["h%d_rewards" % h]
      if is_clip_case(g):  # the wrapper clips at env-step time: restated as oracle.clip_reward per step
        rewards = np.array([oracle.clip_reward(r) for r in rewards.tolist()], np.float64)
      tr, tv, tp = oracle.insert_target(rewards, g["h%d_to_play" % h],
                                        g["h%d_root_values" % h], g["h%d_child_visits" % h], K, T,
                                        disc, int(steps[row]))
      assert np.array_equal(tr, g["b%d_t_rewards" % b][row])
      assert np.array_equal(tp, g["b%d_t_policies" % b][row])
      want = g["b%d_t_values" % b][row]
      # tolerance (north_star): 1e-5 relative in fp32; np.dot's float32 summation order is BLAS's
      scale = np.maximum(np.abs(want), 1.0)
      err = np.max(np.abs(tv - want) / scale)
      worst = max(worst, err)
      assert err <= 1e-5, (b, row, tv, want)
      obs = g["h%d_obs" % h][int(steps[row])].astype(np.float32)
      assert np.array_equal(obs, g["b%d_obs" % b][row])
  print("worst value-target rel err", worst)


def test_transforms():
  g = load("transforms")
  # float32 op order of config.py:53 restated with correctly rounded sqrtf; torch's vectorised CPU
  # sqrt is NOT correctly rounded (~0.7% of inputs are 1 ulp off), so this is a few-ulp tolerance
  h = oracle.scalar_transform(g["x"])
  assert np.mean(h == g["h"]) > 0.98
  assert np.max(np.abs(h - g["h"]) / np.maximum(np.abs(g["h"]), 1e-3)) < 1e-6
  sup = oracle.scalar_to_support(g["support_in"][:, 0], -15, 15)
  assert np.array_equal(sup, g["support"][:, 0, :])
  # h^-1 (config.py:32) cancels catastrophically in float32: a 1-ulp sqrt difference moves the
  # result by ~6e-5 relative.  Same op order => bit-identical except where torch's sqrt is 1 ulp off.
  hi = oracle.inverse_scalar_transform(g["hinv_in"])
  assert np.mean(hi == g["hinv_out"]) > 0.98
  assert np.max(np.abs(hi - g["hinv_out"]) / np.maximum(np.abs(g["hinv_out"]), 1.0)) < 5e-4
  inv_nt = oracle.inverse_transform(g["logits"], -15, 15, True)
  assert np.allclose(inv_nt, g["inverse_no_transform"], rtol=1e-5, atol=1e-6)
  inv = oracle.inverse_transform(g["logits"], -15, 15, False)
  # h^-1 in float32 quantises its output in steps of ~1e-4 (SURVEY.md section 7 hard part 4): a 1-ulp
  # difference in the softmax expectation can move the result by one such step.
  rel = np.abs(inv - g["inverse"]) / np.maximum(np.abs(g["inverse"]), 1.0)
  assert np.mean(rel <= 1e-5) > 0.97
  assert rel.max() < 5e-4
  # the binary64 restatement bounds the error of BOTH float32 pipelines (torch's golden, the float32 oracle)
  truth = oracle.inverse_transform_f64(g["logits"], -15, 15)
  scale = np.maximum(np.abs(truth), 1.0)
  assert np.max(np.abs(g["inverse"].reshape(-1) - truth) / scale) < 5e-4
  assert np.max(np.abs(inv.reshape(-1) - truth) / scale) < 5e-4
  assert np.allclose(oracle.inverse_transform_f64(g["logits"], -15, 15, True), g["inverse_no_transform"].reshape(-1),
                     rtol=1e-5, atol=1e-6)


# ---- the pure-Python port (oracle/search_ref.py, oracle/fcnet_ref.py) used as the CPU baseline ----
@pytest.mark.parametrize("case", ["ttt", "lunar", "atari18"])
def test_python_port_matches_reference_golden(case):
  import torch
  from oracle.search_ref import FlatSearch
  from model_based_rl_b200.testing import HashNetwork
  from helpers import search_case_cfg
  g = load("search_" + case)
  cfg = search_case_cfg(g)
  hv = g["hashnet"]
  net = HashNetwork(cfg["action_space"], float(hv[0]), float(hv[1]), float(hv[2]), int(hv[3]))
  for gi in range(4):
    fs = FlatSearch(**cfg)
    init = net.initial_inference(torch.tensor([[int(g["root_state"][gi])]], dtype=torch.int64))
    legal = [a for a in range(cfg["action_space"]) if (int(g["legal"][gi]) >> a) & 1]
    noise = g["noise"][gi, :len(legal)] if int(g["use_noise"]) else None
    fs.setup_root(init, int(g["to_play"][gi]), legal, noise, float(g["noise_frac"]))
    fs.run(net)
    assert fs.root_visits() == list(g["visits"][gi])
    assert fs.root_value() == g["root_value"][gi]
    assert fs.minmax == tuple(g["minmax"][gi])
    assert [t[0] for t in fs.trace] == list(g["trace_parent"][gi])
    assert [t[1] for t in fs.trace] == list(g["trace_action"][gi])
    assert fs.child_visits() == list(g["visits"][gi] / g["visits"][gi].sum())


def test_python_port_select_action():
  from oracle.search_ref import select_action
  g = load("select_action")
  for i in range(len(g["A"])):
    A, mask = int(g["A"][i]), int(g["legal"][i])
    legal = [a for a in range(A) if (mask >> a) & 1]
    assert select_action(list(g["visits"][i]), legal, g["temperature"][i], g["u"][i]) == g["action"][i]


@pytest.mark.parametrize("name,obs_dim,A", [("atari18", 128, 18), ("ttt", 9, 9)])
def test_fcnet_ref_matches_reference_golden(name, obs_dim, A):
  import torch
  from oracle.fcnet_ref import FCNetworkRef
  g = load("fcnet_" + name)
  net = FCNetworkRef(obs_dim, A)
  net.load_state_dict({k[2:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("w_")})
  with torch.inference_mode():
    init = net.initial_inference(torch.from_numpy(g["obs"]))
    rec = net.recurrent_inference(torch.from_numpy(g["init_hidden"]), g["actions"].tolist())
  for got, want in ((init.hidden_state, "init_hidden"), (init.policy_logits, "init_logits"),
                    (init.value, "init_value"), (rec.hidden_state, "rec_hidden"),
                    (rec.policy_logits, "rec_logits"), (rec.value, "rec_value"),
                    (rec.reward, "rec_reward")):
    assert np.allclose(got.numpy(), g[want], rtol=1e-6, atol=1e-6), want


# ---- prioritized-replay index (oracle/replay_ref.py) against the reference's SumTree --------------
@pytest.mark.parametrize("case", REPLAY_CASES)
def test_replay_index_oracle_matches_reference_golden(case):
  from oracle import replay_ref
  g = load("replay_" + case)
  tree = replay_ref.SumTreeRef(int(g["window_size"]), int(g["window_step"]))
  eps, alpha = float(g["epsilon"]), float(g["alpha"])
  for h in range(int(g["n_hist"])):
    errors = g["h%d_errors" % h].tolist()
    if g["ignores"][h] >= 0:
      errors = errors[:-int(g["ignores"][h])]
    tree.add(replay_ref.get_priorities(errors, eps, alpha) if errors else [], h)
  for b in range(int(g["n_batches"])):
    assert tree.total_priority == g["b%d_total_priority" % b]
    assert tree.num_memories == int(g["b%d_num_memories" % b])
    picks, w = replay_ref.sample_indices(tree, int(g["batch_size"]), g["b%d_frac" % b],
                                         float(g["b%d_beta_after" % b]))
    assert [p[0] for p in picks] == g["b%d_idxs" % b].tolist()
    assert [p[2] for p in picks] == g["b%d_steps" % b].tolist()
    assert [p[3] for p in picks] == g["b%d_hist" % b].tolist()
    assert np.array_equal(np.array([p[1] for p in picks]), g["b%d_priorities" % b])
    assert np.array_equal(w, g["b%d_is_weights" % b])
    pri = replay_ref.get_priorities(g["b%d_update_errors" % b], eps, alpha)
    for idx, p in zip(g["b%d_idxs" % b], pri):
      tree.update(int(idx), p)
    assert np.array_equal(tree.tree, g["b%d_tree_after_update" % b])


@pytest.mark.parametrize("case", REPLAY_CASES)
def test_ring_cursor_matches_reference_golden(case):
  """Host ring arithmetic of the product facade (no GPU needed) vs the reference's num_memories and
  sampled slots."""
  from oracle import replay_ref
  from model_based_rl_b200.replay_buffer import RingCursor
  g = load("replay_" + case)
  ring = RingCursor(int(g["window_size"]), int(g["window_step"]))
  ref = replay_ref.SumTreeRef(int(g["window_size"]), int(g["window_step"]))
  for h in range(int(g["n_hist"])):
    n = len(g["h%d_errors" % h]) - max(0, int(g["ignores"][h]))
    n = max(n, 0)
    before = ref.position
    ref.add([1.0] * n, h)
    slots = ring.take(n)
    if n:
      assert slots[0] == before
    assert (ring.position, ring.capacity, ring.prev_capacity, ring.num_memories) == \
        (ref.position, ref.capacity, ref.prev_capacity, ref.num_memories)
  assert ring.num_memories == int(g["b0_num_memories"])


# ---- MuZeroNetwork (oracle/muzero_ref.py) against the reference class in eval mode --------------
def test_muzero_ref_matches_reference_golden():
  import torch
  from oracle import muzero_ref as mr
  g = load("muzero_net")
  C_in, A = int(g["input_channels"]), int(g["action_space"])
  sd = mr.seeded_state_dict(C_in, A, int(g["seed"]))
  torch.set_num_threads(max(1, torch.get_num_threads()))
  with torch.inference_mode():
    v0, pol0, h0 = mr.initial_inference(torch.from_numpy(g["obs"]), sd)
    v1, r1, pol1, h1 = mr.recurrent_inference(h0, g["actions"].tolist(), sd, A)
    v2, r2, pol2, h2 = mr.recurrent_inference(h1, g["actions2"].tolist(), sd, A)
  for got, want in ((h0, "init_hidden"), (pol0, "init_logits"), (v0, "init_value"), (h1, "rec_hidden"),
                    (pol1, "rec_logits"), (v1, "rec_value"), (r1, "rec_reward"), (h2, "rec2_hidden"),
                    (pol2, "rec2_logits"), (v2, "rec2_value"), (r2, "rec2_reward")):
    # same float32 operators in a different composition (functional vs nn.Module): 1e-4 relative
    assert np.allclose(got.numpy(), g[want], rtol=1e-4, atol=1e-4), want
