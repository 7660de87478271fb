"""Batched self-play driver (model_based_rl_b200/selfplay.py, environments.py) against a fixture
produced by the reference's own Game / TicTacToe / MCTS / select_action (tests/golden/make_golden.py,
gen_selfplay)."""
import types

import numpy as np
import pytest

from helpers import load


def _moves_of(g, game):
  idx = np.nonzero(g["m_game"] == game)[0]
  return idx[np.argsort(g["m_move"][idx])]


def test_vector_tictactoe_matches_reference_env():
  """Host environment: observations, rewards, terminal flags, legal moves and results."""
  from model_based_rl_b200.environments import RESULTS, VectorTicTacToe
  g = load("selfplay_ttt")
  n_games = int(g["n_games"])
  env = VectorTicTacToe(n_games)
  first = env.reset()
  alive = np.ones(n_games, bool)
  cursor = [list(_moves_of(g, i)) for i in range(n_games)]
  for i in range(n_games):
    assert np.array_equal(first[i], g["m_obs"][cursor[i][0]])
  while alive.any():
    rows = [cursor[i][0] if alive[i] else None for i in range(n_games)]
    legal = env.legal_mask()
    actions = np.zeros(n_games, np.int64)
    for i, r in enumerate(rows):
      if r is None:  # a finished slot keeps playing legal filler moves on a fresh board
        actions[i] = int(env.legal_actions(i)[0])
      else:
        assert int(legal[i]) == int(g["m_legal"][r])
        actions[i] = int(g["m_action"][r])
    obs, reward, done, result = env.step(actions)
    for i, r in enumerate(rows):
      if r is None:
        if done[i]:
          env.reset([i])
        continue
      assert bool(done[i]) == bool(g["m_done"][r]) and int(reward[i]) == int(g["m_reward"][r])
      cursor[i].pop(0)
      if done[i]:
        assert RESULTS[int(result[i])] == str(g["g_result"][i])
        assert int(env.elapsed[i]) == int(g["g_length"][i])
        alive[i] = False
        env.reset([i])
      else:
        assert np.array_equal(obs[i].astype(np.float32), g["m_next_obs"][r])


@pytest.mark.gpu
def test_batched_actor_replays_reference_selfplay():
  """Eight games in lock step on the GPU reproduce the reference's sequential games move for move:
  actions, root values, child-visit distributions, priority seeds, and the history slices handed to
  the replay buffer (chunking at max_history_length, overlap, ignore, terminal)."""
  import torch
  from model_based_rl_b200.environments import VectorTicTacToe
  from model_based_rl_b200.selfplay import BatchedActor
  from model_based_rl_b200.testing import ObsHashNetwork
  g = load("selfplay_ttt")
  n_games, A, S = int(g["n_games"]), int(g["A"]), int(g["S"])
  hv = g["hashnet"]
  cfg = types.SimpleNamespace(
      num_simulations=S, action_space=A, two_players=True, discount=1.0, pb_c_base=19652, pb_c_init=1.25,
      init_value_score=0.0, known_bounds=[-1, 1], root_dirichlet_alpha=0.25, root_exploration_fraction=0.25,
      num_unroll_steps=int(g["K"]), td_steps=int(g["T"]), max_history_length=int(g["max_history_length"]),
      max_steps=10 ** 9)
  net = ObsHashNetwork(A, float(hv[0]), float(hv[1]), float(hv[2]), int(hv[3]), device="cuda")
  temps = np.array([[1.0, 0.5, 0.0, 0.25][i % 4] for i in range(n_games)])
  actor = BatchedActor(cfg, net, VectorTicTacToe(n_games), device="cuda", temperature=temps)
  cursor = [list(_moves_of(g, i)) for i in range(n_games)]
  rng = np.random.default_rng(0)
  checked = 0
  while any(cursor):
    noise = rng.dirichlet([0.25] * A, size=n_games)
    u = rng.random(n_games)
    rows = [c[0] if c else None for c in cursor]
    for i, r in enumerate(rows):
      if r is not None:
        noise[i], u[i] = g["m_noise"][r], g["m_u"][r]
    actions, root_value, child_visits, errors, done = actor.play_move(noise, u)
    for i, r in enumerate(rows):
      if r is None:
        continue
      assert int(actions[i]) == int(g["m_action"][r]), (i, r)
      assert root_value[i] == g["m_root_value"][r]                      # bit-exact search
      assert np.array_equal(child_visits[i], g["m_child_visits"][r])
      assert errors[i] == g["m_error"][r]
      assert bool(done[i]) == bool(g["m_done"][r])
      cursor[i].pop(0)
      checked += 1
  assert checked == len(g["m_game"])
  # history hand-off: the first game of every slot, in order
  for i in range(n_games):
    want = np.nonzero(g["s_game"] == i)[0]
    got = [s for s in actor.saved if s[0] == i][:len(want)]
    assert len(got) == len(want)
    for (_, h, ignore, terminal), w in zip(got, want):
      assert len(h.root_values) == int(g["s_n"][w]) and len(h.observations) == int(g["s_n_obs"][w])
      assert (-1 if ignore is None else ignore) == int(g["s_ignore"][w])
      assert bool(terminal) == bool(g["s_terminal"][w])
      assert ";".join(map(str, h.actions)) == str(g["s_actions"][w])
  assert actor.games_played >= n_games and sum(actor.results.values()) >= n_games


@pytest.mark.gpu
def test_actor_feeds_prioritized_replay():
  """End to end: batched self-play -> PrioritizedReplay (device window + sum-tree) -> sampled batch
  whose targets match the oracle's insert_target for the history each row came from."""
  import oracle
  from model_based_rl_b200.environments import VectorTicTacToe
  from model_based_rl_b200.replay_buffer import PrioritizedReplay
  from model_based_rl_b200.selfplay import BatchedActor
  from model_based_rl_b200.testing import ObsHashNetwork
  A, G = 9, 64
  cfg = types.SimpleNamespace(
      num_simulations=10, action_space=A, two_players=True, discount=1.0, pb_c_base=19652, pb_c_init=1.25,
      init_value_score=0.0, known_bounds=[-1, 1], root_dirichlet_alpha=0.25, root_exploration_fraction=0.25,
      num_unroll_steps=3, td_steps=4, max_history_length=500, max_steps=10 ** 9, batch_size=32,
      beta_increment_per_sampling=0.001, epsilon=0.01, alpha=1.0, beta=0.4, obs_space=(9,), window_size=2048,
      window_step=None, seed=3)
  rb = PrioritizedReplay(cfg)
  kept = []
  real_save = rb.save_history
  rb.save_history = lambda h, ignore=None, terminal=False: (kept.append(h), real_save(h, ignore, terminal))[1]
  net = ObsHashNetwork(A, 0.8, 0.5, 2.0, 3, device="cuda")
  actor = BatchedActor(cfg, net, VectorTicTacToe(G), replay_buffer=rb, device="cuda")
  np.random.seed(1)
  for _ in range(12):
    actor.play_move()
  assert actor.games_played >= G and rb.size() == sum(len(h.root_values) for h in kept)
  (obs, actions, (t_r, t_v, t_p)), idxs, is_w = rb.sample_batch()
  assert obs.shape == (32, 9) and t_p.shape == (32, 4, A) and len(actions) == 32 and is_w.max() == 1.0
  for b in range(32):  # locate the (history, step) of the row through its slot and check the targets
    slot = idxs[b] - (2048 - 1)
    h = kept[int(rb.index.slot_chunk[slot])]
    steps = [s for s in range(len(h.root_values)) if np.array_equal(np.float32(h.observations[s]), obs[b])]
    ok = False
    for step in steps:
      r, v, p = oracle.insert_target(np.array(h.rewards, np.float64), np.array(h.to_play, np.int8),
                                     np.array(h.root_values), np.array(h.child_visits), 3, 4, 1.0, step)
      ok |= bool(np.array_equal(t_r[b], r) and np.array_equal(t_p[b], p) and np.allclose(t_v[b], v, rtol=1e-5, atol=1e-6))
    assert ok, b


@pytest.mark.gpu
def test_fast_path_through_fcsearch_equals_generic_path():
  """BatchedActor(search=FCSearch): the whole move behind one search_host call (CUDA graph, legal masks
  and to_play staged from the host) plays exactly the games the generic per-simulation path plays with
  the same FCNetwork (float32 kernels on both sides): actions, root values, child visits, priority
  seeds and the histories handed over, bit for bit."""
  import torch
  from model_based_rl_b200.environments import VectorTicTacToe
  from model_based_rl_b200.networks import FCNetwork, FCSearch, random_state_dict
  from model_based_rl_b200.selfplay import BatchedActor
  A, G, S = 9, 96, 25
  cfg = types.SimpleNamespace(
      num_simulations=S, action_space=A, two_players=True, discount=1.0, pb_c_base=19652, pb_c_init=1.25,
      init_value_score=0.0, known_bounds=[-1, 1], root_dirichlet_alpha=0.25, root_exploration_fraction=0.25,
      num_unroll_steps=3, td_steps=4, max_history_length=500, max_steps=10 ** 9,
      value_support=[-15, 15], reward_support=[-15, 15], no_support=False, no_target_transform=False)
  net = FCNetwork(9, A, "cuda", cfg, precision="f32")
  net.load_weights(random_state_dict(9, A, seed=7))
  temps = np.array([[1.0, 0.5, 0.0, 0.25][i % 4] for i in range(G)])
  slow = BatchedActor(cfg, net, VectorTicTacToe(G), device="cuda", temperature=temps)
  fast = BatchedActor(cfg, net, VectorTicTacToe(G), device="cuda", temperature=temps,
                      search=FCSearch(cfg, net, G, use_graph=True, num_streams=2))
  rng = np.random.default_rng(11)
  for move in range(14):
    noise = np.zeros((G, A))
    legal = slow.env.legal_mask()
    assert np.array_equal(legal, fast.env.legal_mask())
    for i in range(G):
      n = bin(int(legal[i])).count("1")
      noise[i, :n] = rng.dirichlet([0.25] * n)
    u = rng.random(G)
    a = slow.play_move(noise.copy(), u.copy())
    b = fast.play_move(noise.copy(), u.copy())
    for x, y, name in zip(a, b, ("actions", "root_value", "child_visits", "errors", "done")):
      assert np.array_equal(np.asarray(x), np.asarray(y)), (move, name)
  assert fast.games_played == slow.games_played >= G
  assert len(fast.saved) == len(slow.saved)
  for (i, h, ign, term), (j, h2, ign2, term2) in zip(fast.saved, slow.saved):
    assert (i, ign, term) == (j, ign2, term2) and h.actions == h2.actions and h.root_values == h2.root_values
    assert h.child_visits == h2.child_visits and h.errors == h2.errors and h.to_play == h2.to_play


def test_synthetic_frames_environment_contract():
  """environments.SyntheticFrames: the vector-environment contract BatchedActor relies on (reset / step /
  legal_mask / elapsed), deterministic under its seed."""
  from model_based_rl_b200.environments import SyntheticFrames
  a, b = SyntheticFrames(3, 6, (2, 8, 8), episode_length=2, seed=4), SyntheticFrames(3, 6, (2, 8, 8), episode_length=2, seed=4)
  o1, o2 = a.reset(), b.reset()
  assert o1.shape == (3, 2, 8, 8) and o1.dtype == np.float32 and np.array_equal(o1, o2)
  assert (a.legal_mask() == 63).all()
  obs, reward, done, result = a.step(np.array([0, 5, 3]))
  obs2, reward2, done2, _ = b.step(np.array([1, 1, 1]))
  assert np.array_equal(obs, obs2) and np.array_equal(reward, reward2)  # frames do not depend on the actions
  assert set(np.unique(reward)) <= {-1, 0, 1} and not done.any() and (result == -1).all()
  _, _, done, _ = a.step(np.array([0, 0, 0]))
  assert done.all() and (a.elapsed == 2).all()
  assert a.reset([1]).shape == (1, 2, 8, 8) and a.elapsed.tolist() == [2, 0, 2]
  with pytest.raises(ValueError):
    a.step(np.array([0, 6, 0]))


@pytest.mark.gpu
@pytest.mark.parametrize("env_name", ["tictactoe", "ram"])
def test_device_actor_fills_the_replay_like_the_list_path(env_name):
  """DeviceActor (trajectories written on the device, straight from the search engine's output buffers into the
  replay window) against BatchedActor + save_history (per-game Python lists -> HistorySlice -> upload): same
  moves, and afterwards the same replay buffer -- sum-tree contents, num_memories, and for every slot the same
  (step, chunk length) and the same chunk contents (observations, actions, rewards, to_play, root values,
  child visits), bit for bit.  Short max_history_length so that chunks of running games overlap
  (actors.py:160-169), byte observations with device-side normalisation in the `ram` case."""
  from model_based_rl_b200.environments import SyntheticRam, VectorTicTacToe
  from model_based_rl_b200.networks import FCNetwork, FCSearch, random_state_dict
  from model_based_rl_b200.replay_buffer import PrioritizedReplay
  from model_based_rl_b200.selfplay import BatchedActor, DeviceActor
  import torch
  if env_name == "tictactoe":
    A, G, S, D, two, L, moves = 9, 64, 12, 9, True, 4, 30
    make_env = lambda: VectorTicTacToe(G)
  else:
    A, G, S, D, two, L, moves = 6, 96, 10, 128, False, 5, 40
    make_env = lambda: SyntheticRam(G, A, D, episode_length=17, seed=5)
  cfg = types.SimpleNamespace(
      num_simulations=S, action_space=A, two_players=two, discount=1.0 if two else 0.997, pb_c_base=19652, pb_c_init=1.25,
      init_value_score=0.0, known_bounds=[-1, 1] if two else [None, None], root_dirichlet_alpha=0.25,
      root_exploration_fraction=0.25, num_unroll_steps=3, td_steps=4, max_history_length=L, max_steps=10 ** 9,
      value_support=[-15, 15], reward_support=[-15, 15], no_support=False, no_target_transform=False,
      batch_size=32, beta_increment_per_sampling=0.001, epsilon=0.01, alpha=0.8, beta=0.5, obs_space=(D,),
      window_size=600, window_step=None, seed=None, clip_rewards=False)
  net = FCNetwork(D, A, "cuda", cfg)
  net.load_weights(random_state_dict(D, A, seed=3))
  temps = np.array([[1.0, 0.5, 0.0, 0.25][i % 4] for i in range(G)])
  rb_a = PrioritizedReplay(cfg, window_positions=40000)
  rb_b = PrioritizedReplay(cfg, window_positions=40000)
  fs_a, fs_b = FCSearch(cfg, net, G), FCSearch(cfg, net, G)
  # the list path feeds float32(bytes) to the network; the same values on the device: (x - 0) / 1
  fs_b.set_obs_normalization(np.zeros(D, np.float32), np.ones(D, np.float32))
  list_actor = BatchedActor(cfg, net, make_env(), replay_buffer=rb_a, device="cuda", temperature=temps, search=fs_a)
  dev_actor = DeviceActor(cfg, make_env(), rb_b, fs_b, temperature=temps)
  rng = np.random.default_rng(17)
  for move in range(moves):
    legal = list_actor.env.legal_mask()
    assert np.array_equal(legal, dev_actor.env.legal_mask())
    noise = np.zeros((G, A))
    for i in range(G):
      n = bin(int(legal[i])).count("1")
      noise[i, :n] = rng.dirichlet([0.25] * n)
    u = rng.random(G)
    a = list_actor.play_move(noise.copy(), u.copy())
    b = dev_actor.play_move(noise.copy(), u.copy())
    for x, y, name in zip(a, b, ("actions", "root_value", "child_visits", "errors", "done")):
      assert np.array_equal(np.asarray(x), np.asarray(y)), (move, name)
  torch.cuda.synchronize()
  assert dev_actor.games_played == list_actor.games_played > 0
  assert rb_a.size() == rb_b.size() > 0
  assert rb_a.get_throughput() == rb_b.get_throughput()
  assert np.array_equal(rb_a.index.tree.cpu().numpy(), rb_b.index.tree.cpu().numpy())
  ia, ib = rb_a.index, rb_b.index
  pa, sa, la = ia.slot_pos.cpu().numpy(), ia.slot_start.cpu().numpy(), ia.slot_len.cpu().numpy()
  pb, sb, lb = ib.slot_pos.cpu().numpy(), ib.slot_start.cpu().numpy(), ib.slot_len.cpu().numpy()
  live = ia.slot_chunk >= 0
  assert np.array_equal(live, ib.slot_chunk >= 0) and live.sum() == min(600, rb_a.size())
  assert np.array_equal((pa - sa)[live], (pb - sb)[live]) and np.array_equal(la[live], lb[live])
  wa = {k: getattr(rb_a, k).cpu().numpy() for k in ("w_obs", "w_actions", "w_rewards", "w_to_play", "w_root_values", "w_child_visits")}
  wb = {k: getattr(rb_b, k).cpu().numpy() for k in wa}
  checked = set()
  for slot in np.nonzero(live)[0]:
    key = (int(sa[slot]), int(sb[slot]))
    if key in checked:
      continue
    checked.add(key)
    n = int(la[slot])
    for k in wa:
      assert np.array_equal(wa[k][sa[slot]:sa[slot] + n], wb[k][sb[slot]:sb[slot] + n]), (k, slot)
  assert len(checked) > 10


@pytest.mark.gpu
def test_device_dirichlet_noise_has_the_reference_distribution():
  """mz_dirichlet_noise against np.random.dirichlet([alpha] * n) (mcts.py:59): rows are dense over the legal actions
  and sum to one, every component follows Beta(alpha, (n - 1) * alpha) (Kolmogorov-Smirnov against scipy's CDF and
  against numpy's own draws), components of a row are negatively correlated by -1 / (n - 1), draws repeat for equal
  (seed, move) and differ otherwise."""
  import ctypes as C
  import torch
  from scipy import stats
  from model_based_rl_b200 import _lib
  lib = _lib.load()
  G, A, alpha = 20000, 18, 0.25
  rng = np.random.default_rng(0)
  legal = np.full(G, (1 << A) - 1, np.int64)
  legal[G // 2:] = rng.integers(1, 1 << A, size=G - G // 2)
  d_legal = torch.from_numpy(legal.astype(np.int32)).cuda()
  out = torch.full((G, A), -1.0, dtype=torch.float64, device="cuda")
  def draw(seed, move):
    _lib.check(lib.mz_dirichlet_noise(G, A, alpha, _lib.ptr(d_legal), seed, move, _lib.ptr(out), None), "noise")
    torch.cuda.synchronize()
    return out.cpu().numpy().copy()
  x = draw(7, 0)
  n = np.array([bin(int(m)).count("1") for m in legal])
  assert np.allclose(x.sum(1), 1.0, atol=1e-12) and (x >= 0).all()
  assert all((x[g, n[g]:] == 0).all() for g in range(G // 2, G, 97))
  full = x[:G // 2]
  for col in (0, 7, 17):
    assert stats.kstest(full[:, col], stats.beta(alpha, (A - 1) * alpha).cdf).pvalue > 1e-3
    assert stats.ks_2samp(full[:, col], rng.dirichlet([alpha] * A, size=G // 2)[:, col]).pvalue > 1e-3
  corr = np.corrcoef(full[:, 0], full[:, 1])[0, 1]
  assert abs(corr - (-1.0 / (A - 1))) < 0.03
  # a subset of legal actions: components over n_g children follow Beta(alpha, (n_g - 1) * alpha)
  for k in (2, 5, 9):
    rows = np.nonzero((n == k) & (np.arange(G) >= G // 2))[0]
    if len(rows) > 300:
      assert stats.kstest(x[rows, 0], stats.beta(alpha, (k - 1) * alpha).cdf).pvalue > 1e-3
  assert (x[n == 1, 0] == 1.0).all()
  assert np.array_equal(draw(7, 0), x)
  assert not np.array_equal(draw(7, 1), x) and not np.array_equal(draw(8, 0), x)
  # alpha >= 1 takes the Marsaglia-Tsang branch without the power correction
  alpha = 1.5
  y = draw(3, 0)[:G // 2]
  assert stats.kstest(y[:, 4], stats.beta(alpha, (A - 1) * alpha).cdf).pvalue > 1e-3


@pytest.mark.gpu
def test_device_actor_with_device_noise_runs_the_whole_loop():
  """DeviceActor.play_move with nothing brought by the caller: Dirichlet noise drawn on the device, uniforms on the
  host; moves are legal, chunk commits go through commit_chunks, sampling from the filled replay buffer works."""
  from model_based_rl_b200.environments import SyntheticRam
  from model_based_rl_b200.networks import FCNetwork, FCSearch, random_state_dict
  from model_based_rl_b200.replay_buffer import PrioritizedReplay
  from model_based_rl_b200.selfplay import DeviceActor
  import torch
  A, G, S, D = 6, 256, 8, 128
  cfg = types.SimpleNamespace(
      num_simulations=S, action_space=A, two_players=False, discount=0.997, pb_c_base=19652, pb_c_init=1.25,
      init_value_score=0.0, known_bounds=[None, None], root_dirichlet_alpha=0.25, root_exploration_fraction=0.25,
      num_unroll_steps=3, td_steps=4, max_history_length=20, max_steps=10 ** 9, value_support=[-15, 15],
      reward_support=[-15, 15], no_support=False, no_target_transform=False, batch_size=64,
      beta_increment_per_sampling=0.001, epsilon=0.01, alpha=1.0, beta=0.5, obs_space=(D,), window_size=3000,
      window_step=None, seed=None, clip_rewards=True)
  np.random.seed(4)
  net = FCNetwork(D, A, "cuda", cfg)
  net.load_weights(random_state_dict(D, A, seed=3))
  rb = PrioritizedReplay(cfg, window_positions=60000)
  env = SyntheticRam(G, A, D, episode_length=33, seed=2)
  actor = DeviceActor(cfg, env, rb, FCSearch(cfg, net, G))
  env.elapsed[:] = np.arange(G) % 33  # games end on different moves
  seen = []
  for _ in range(60):
    actions, root_value, child_visits, errors, done = actor.play_move()
    assert ((actions >= 0) & (actions < A)).all()
    assert np.allclose(np.asarray(child_visits).sum(1), 1.0)
    seen.append(actor.search.noise.cpu().numpy().copy())
  assert not np.array_equal(seen[0], seen[1])
  assert np.allclose(seen[-1].sum(1), 1.0)
  assert actor.games_played > G and rb.size() > 2000
  (obs, acts, t_r, t_v, t_p, vs, rs), idx, isw = rb.sample_batch_device(True)
  torch.cuda.synchronize()
  assert obs.shape == (64, D) and torch.isfinite(t_v).all() and float(isw.max()) == 1.0


@pytest.mark.gpu
def test_pipelined_actors_equal_actors_taking_turns():
  """PipelinedActors (every actor's next search is enqueued before the next actor's host work starts) against the same
  two DeviceActors called in turn: same actions, root values and errors move by move, same replay buffer at the end
  (size, sum-tree total, a sampled batch) -- the overlap changes when things run, not what is computed."""
  import random
  from model_based_rl_b200.environments import SyntheticRam
  from model_based_rl_b200.networks import FCNetwork, FCSearch, random_state_dict
  from model_based_rl_b200.replay_buffer import PrioritizedReplay
  from model_based_rl_b200.selfplay import DeviceActor, PipelinedActors
  import torch
  A, G, S, D = 6, 192, 8, 128
  cfg = types.SimpleNamespace(
      num_simulations=S, action_space=A, two_players=False, discount=0.997, pb_c_base=19652, pb_c_init=1.25,
      init_value_score=0.0, known_bounds=[None, None], root_dirichlet_alpha=0.25, root_exploration_fraction=0.25,
      num_unroll_steps=3, td_steps=4, max_history_length=20, max_steps=10 ** 9, value_support=[-15, 15],
      reward_support=[-15, 15], no_support=False, no_target_transform=False, batch_size=64,
      beta_increment_per_sampling=0.001, epsilon=0.01, alpha=1.0, beta=0.5, obs_space=(D,), window_size=4000,
      window_step=None, seed=None, clip_rewards=True)
  net = FCNetwork(D, A, "cuda", cfg)
  net.load_weights(random_state_dict(D, A, seed=3))

  def build():
    np.random.seed(11)
    random.seed(11)
    rb = PrioritizedReplay(cfg, window_positions=80000)
    actors = []
    for k in range(2):
      env = SyntheticRam(G, A, D, episode_length=29 + 4 * k, seed=5 + k)
      actors.append(DeviceActor(cfg, env, rb, FCSearch(cfg, net, G)))
      env.elapsed[:] = np.arange(G) % (29 + 4 * k)  # games end on different moves
    return rb, actors

  moves = 45
  rb_a, turn = build()
  log_a = []
  for _ in range(moves + 1):  # one more: the pipelined run ends with a move of every actor in flight
    for a in turn:
      r = a.play_move()
      log_a.append((r[0].copy(), r[1].numpy().copy(), r[3].copy(), r[4].copy()))
  rb_b, piped = build()
  pipe = PipelinedActors(piped, copy_outputs=True)
  log_b = []
  for _ in range(moves):
    for r in pipe.play_round():
      log_b.append((r[0].copy(), r[1].numpy().copy(), r[3].copy(), r[4].copy()))
  for r in pipe.drain():  # the moves in flight
    log_b.append((r[0].copy(), r[1].numpy().copy(), r[3].copy(), r[4].copy()))
  assert len(log_a) == len(log_b) == 2 * (moves + 1)
  for x, y in zip(log_a, log_b):
    for u, v in zip(x, y):
      assert np.array_equal(u, v)
  torch.cuda.synchronize()
  assert rb_a.size() == rb_b.size() and rb_a.size() > 2000
  assert rb_a.index.total_priority == rb_b.index.total_priority
  assert sum(a.games_played for a in turn) == sum(a.games_played for a in piped)
  random.seed(1), np.random.seed(1)
  sa = rb_a.sample_batch()
  random.seed(1), np.random.seed(1)
  sb = rb_b.sample_batch()
  assert sa[1] == sb[1] and np.array_equal(sa[2], sb[2])
  for u, v in zip(sa[0][2], sb[0][2]):
    assert np.array_equal(u, v)
  assert sa[0][1] == sb[0][1]
  assert np.array_equal(sa[0][0], sb[0][0]), "observation rows differ: %s" % np.nonzero((sa[0][0] != sb[0][0]).any(1))[0]
