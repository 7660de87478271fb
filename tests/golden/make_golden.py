"""Generates tests/golden/*.npz by running the UNMODIFIED reference modules.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

The reference ships no tests or golden vectors, so parity is pinned on outputs
of its own code: mcts.MCTS / mcts.Node / mcts.MinMaxStats, config.Config (select_action and the
transforms), game.Game.store_search_statistics, replay_buffer.PrioritizedReplay
(save_history / sample_batch / insert_target / update, with a 3-line stub for the `ray` decorator)
and networks.FCNetwork.  The only things replaced are the random sources (np.random.dirichlet,
np.random.choice, random.uniform, np.random.randint), which are fed from recorded buffers so that
the oracle and the CUDA engine can consume identical numbers, and the network, which is the
batch-invariant HashNetwork for the search fixtures.
"""
import math
import os
import random
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("MZ_REFERENCE", "/root/reference")
sys.path.insert(0, REPO)
sys.path.insert(0, REF)

ray_stub = types.ModuleType("ray")
ray_stub.remote = lambda cls: cls
sys.modules.setdefault("ray", ray_stub)

import mcts as ref_mcts  # noqa: E402  (reference)
import config as ref_config  # noqa: E402
import game as ref_game  # noqa: E402
import replay_buffer as ref_replay  # noqa: E402
import networks as ref_networks  # noqa: E402

from model_based_rl_b200.testing import HashNetwork  # noqa: E402

BASE_CFG = dict(
    value_support=[-15, 15], reward_support=[-15, 15], no_support=False, no_target_transform=False,
    num_simulations=30, discount=0.997, pb_c_base=19652, pb_c_init=1.25, init_value_score=0.0,
    action_space=4, two_players=False, known_bounds=[None, None], root_dirichlet_alpha=0.25,
    root_exploration_fraction=0.25, episode_life=False, clip_rewards=False, sticky_actions=1,
    batch_size=32, beta_increment_per_sampling=0.001, epsilon=0.01, alpha=1.0, beta=1.0,
    stored_before_train=0, num_unroll_steps=5, td_steps=10, obs_space=(8,), window_size=4096,
    window_step=None, seed=None)


def make_config(**kw):
  d = dict(BASE_CFG)
  d.update(kw)
  return ref_config.Config(d)


# --------------------------------------------------------------------------------------------
# search
# --------------------------------------------------------------------------------------------
def run_reference_search(cfg, net, root_state, legal, noise, to_play):
  """Restates the move body of Actor.play_game (actors.py:131-147) around the reference classes."""
  G = len(root_state)
  S, A = cfg.num_simulations, cfg.action_space
  engine = ref_mcts.MCTS(cfg)
  out = dict(
      visits=np.zeros((G, A), np.int32), root_value=np.zeros(G), root_vsum=np.zeros(G),
      minmax=np.zeros((G, 2)), trace_parent=np.zeros((G, S), np.int32),
      trace_action=np.zeros((G, S), np.int32), trace_depth=np.zeros((G, S), np.int32),
      edge_prior=np.zeros((G, S + 1, A)), edge_vsum=np.zeros((G, S + 1, A)),
      edge_visit=np.zeros((G, S + 1, A), np.int32), edge_reward=np.zeros((G, S + 1, A)),
      edge_child=np.full((G, S + 1, A), -2, np.int32), root_logits=np.zeros((G, A), np.float32),
      root_net_value=np.zeros(G, np.float32))
  real_dirichlet = np.random.dirichlet
  for g in range(G):
    root = ref_mcts.Node(0)
    state = torch.tensor([[int(root_state[g])]], dtype=torch.int64)
    init = net.initial_inference(state)
    out["root_logits"][g] = init.policy_logits[0].numpy()
    out["root_net_value"][g] = init.value.item()
    legal_actions = [a for a in range(A) if (int(legal[g]) >> a) & 1]
    root.expand(init, int(to_play[g]), legal_actions)
    if noise is not None:
      np.random.dirichlet = lambda alpha, g=g: noise[g, :len(alpha)].copy()
      root.add_exploration_noise(cfg.root_dirichlet_alpha, cfg.root_exploration_fraction)
      np.random.dirichlet = real_dirichlet
    paths = engine.run(root, net)

    ids = {id(root): 0}
    nodes = [root]
    for s, path in enumerate(paths):
      leaf = path[-1]
      assert id(leaf) not in ids
      ids[id(leaf)] = s + 1
      nodes.append(leaf)
      parent = path[-2]
      act = [a for a, c in parent.children.items() if c is leaf][0]
      out["trace_parent"][g, s] = ids[id(parent)]
      out["trace_action"][g, s] = act
      out["trace_depth"][g, s] = len(path) - 1
    for n, node in enumerate(nodes):
      for a, child in node.children.items():
        out["edge_prior"][g, n, a] = child.prior
        out["edge_vsum"][g, n, a] = child.value_sum
        out["edge_visit"][g, n, a] = child.visit_count
        out["edge_reward"][g, n, a] = child.reward
        out["edge_child"][g, n, a] = ids.get(id(child), -1)
    for a, child in root.children.items():
      out["visits"][g, a] = child.visit_count
    out["root_value"][g] = root.value()
    out["root_vsum"][g] = root.value_sum
    out["minmax"][g] = (engine.min_max_stats.minimum, engine.min_max_stats.maximum)
  return out


SEARCH_CASES = {
    # name: (config overrides, G, hashnet kwargs, use_noise, legal subsets, random to_play)
    "ttt": (dict(action_space=9, num_simulations=30, two_players=True, discount=1.0,
                 known_bounds=[-1.0, 1.0], td_steps=10),
            12, dict(value_scale=1.0, reward_scale=1.0, logit_scale=2.0, reward_density=2), True,
            True, True),
    "lunar": (dict(action_space=4, num_simulations=30, discount=0.997, td_steps=1000),
              12, dict(value_scale=4.0, reward_scale=2.0, logit_scale=1.5, reward_density=6), True,
              False, False),
    "breakout": (dict(action_space=4, num_simulations=50, discount=0.997),
                 8, dict(value_scale=2.0, reward_scale=1.0, logit_scale=3.0, reward_density=1), True,
                 False, False),
    "atari18": (dict(action_space=18, num_simulations=50, discount=0.997),
                8, dict(value_scale=1.0, reward_scale=0.5, logit_scale=2.0, reward_density=3), True,
                False, False),
    "atari18_nonoise_bounds": (dict(action_space=18, num_simulations=50, discount=0.95,
                                    known_bounds=[-2.0, 2.0], init_value_score=0.5),
                               4, dict(value_scale=1.0, reward_scale=0.5, logit_scale=0.25,
                                       reward_density=8), False, False, False),
    "flat_ties": (dict(action_space=6, num_simulations=20, discount=1.0),
                  4, dict(value_scale=0.0, reward_scale=0.0, logit_scale=0.0, reward_density=0),
                  False, False, False),
}


def gen_search(rng):
  for name, (over, G, hk, use_noise, subsets, rand_tp) in SEARCH_CASES.items():
    cfg = make_config(**over)
    A = cfg.action_space
    net = HashNetwork(A, **hk)
    root_state = rng.integers(0, 2**63 - 1, size=G, dtype=np.int64)
    if subsets:
      legal = rng.integers(1, 2**A, size=G, dtype=np.int64).astype(np.uint32)
      legal[0] = (1 << A) - 1
      legal[1] = 1 << (A - 1)  # a single legal move
    else:
      legal = np.full(G, (1 << A) - 1, np.uint32)
    noise = None
    if use_noise:
      noise = np.zeros((G, A))
      for g in range(G):
        n = bin(int(legal[g])).count("1")
        noise[g, :n] = rng.dirichlet([cfg.root_dirichlet_alpha] * n)
    to_play = rng.choice([-1, 1], size=G).astype(np.int8) if rand_tp else np.ones(G, np.int8)
    out = run_reference_search(cfg, net, root_state, legal, noise, to_play)
    np.savez_compressed(
        os.path.join(HERE, "search_%s.npz" % name), root_state=root_state, legal=legal,
        noise=noise if noise is not None else np.zeros((0, A)), use_noise=np.int32(use_noise),
        to_play=to_play, num_simulations=np.int32(cfg.num_simulations), action_space=np.int32(A),
        two_players=np.int32(cfg.two_players), discount=np.float64(cfg.discount),
        pb_c_base=np.float64(cfg.pb_c_base), pb_c_init=np.float64(cfg.pb_c_init),
        init_value_score=np.float64(cfg.init_value_score),
        known_bounds=np.array([np.nan if b is None else b for b in cfg.known_bounds]),
        noise_frac=np.float64(cfg.root_exploration_fraction),
        hashnet=np.array([hk["value_scale"], hk["reward_scale"], hk["logit_scale"],
                          hk["reward_density"]], np.float64),
        py_sum_mode=np.int32(1 if sys.version_info >= (3, 12) else 0), **out)
    print("search", name, "mean depth %.2f" % out["trace_depth"].mean(),
          "max depth", out["trace_depth"].max())


# --------------------------------------------------------------------------------------------
# select_action / store_search_statistics
# --------------------------------------------------------------------------------------------
class _U(object):
  """np.random.choice stand-in driven by one host-supplied uniform (legacy RandomState.choice
  draws exactly one random_sample() and searchsorts the normalised cdf, side='right')."""

  def __init__(self):
    self.u = None

  def __call__(self, a, p=None):
    if p is not None:
      cdf = np.asarray(p, dtype=np.float64).cumsum()
      cdf /= cdf[-1]
      return int(cdf.searchsorted(self.u, side='right'))
    arr = np.arange(a) if np.isscalar(a) else np.asarray(a)
    return arr[int(math.floor(self.u * len(arr)))]


def gen_select_action(rng):
  # first: the stand-in reproduces numpy's own draw for the same underlying uniform
  for seed in range(200):
    n = int(rng.integers(1, 19))
    p = rng.random(n) + 1e-3
    p /= p.sum()
    np.random.seed(seed)
    u = np.random.random_sample()
    np.random.seed(seed)
    want = np.random.choice(n, p=p)
    fake = _U()
    fake.u = u
    assert fake(n, p=p) == want, (seed, n)
  cfg = make_config()
  fake = _U()
  real_choice = np.random.choice
  rows = []
  np.random.choice = fake
  try:
    for i in range(400):
      A = int(rng.choice([4, 9, 18]))
      n_children = A if i % 3 else int(rng.integers(1, A + 1))
      actions = sorted(rng.choice(A, size=n_children, replace=False).tolist())
      S = int(rng.choice([30, 50]))
      visits = rng.multinomial(S, rng.dirichlet([0.3] * n_children))
      if i % 7 == 0:
        visits[:] = S // n_children  # many ties
      T = float(rng.choice([0.0, 0.1, 0.25, 0.5, 0.7, 1.0]))
      u = float(rng.random())
      root = ref_mcts.Node(0)
      for a, v in zip(actions, visits):
        root.children[a] = ref_mcts.Node(0.1)
        root.children[a].visit_count = int(v)
      root.visit_count = int(visits.sum())
      root.value_sum = float(rng.normal()) * root.visit_count
      fake.u = u
      action = cfg.select_action(root, T)
      # store_search_statistics only touches game.history / running stats
      g = types.SimpleNamespace(history=types.SimpleNamespace(child_visits=[], root_values=[]),
                                action_space=range(A), sum_values=0, max_value=-np.inf)
      ref_game.Game.store_search_statistics(g, root)
      v_full = np.zeros(18, np.int32)
      mask = 0
      for a, v in zip(actions, visits):
        v_full[a] = v
        mask |= 1 << a
      cv = np.zeros(18)
      cv[:A] = g.history.child_visits[0]
      rows.append((A, mask, T, u, int(action), v_full, cv, g.history.root_values[0],
                   root.value_sum, root.visit_count))
  finally:
    np.random.choice = real_choice
  np.savez_compressed(
      os.path.join(HERE, "select_action.npz"),
      A=np.array([r[0] for r in rows], np.int32), legal=np.array([r[1] for r in rows], np.uint32),
      temperature=np.array([r[2] for r in rows]), u=np.array([r[3] for r in rows]),
      action=np.array([r[4] for r in rows], np.int32), visits=np.stack([r[5] for r in rows]),
      child_visits=np.stack([r[6] for r in rows]), root_value=np.array([r[7] for r in rows]),
      root_vsum=np.array([r[8] for r in rows]), root_visit=np.array([r[9] for r in rows], np.int32))
  print("select_action rows", len(rows))


# --------------------------------------------------------------------------------------------
# replay / targets
# --------------------------------------------------------------------------------------------
def synth_history(rng, n, A, obs_dim, two_players, obs_uint8, running):
  """A HistorySlice shaped like what Actor.play_game sends (actors.py:160-169)."""
  n_obs = n + (1 if running else 0)
  if obs_uint8:
    obs = [rng.integers(0, 256, size=obs_dim, dtype=np.uint8) for _ in range(n_obs)]
  else:
    obs = [rng.normal(size=obs_dim).astype(np.float32) for _ in range(n_obs)]
  cv = []
  for _ in range(n):
    c = rng.multinomial(30, rng.dirichlet([0.5] * A))
    cv.append([int(x) / 30 for x in c])
  root_values = [float(x) for x in rng.normal(0, 2, size=n)]
  actions = [int(x) for x in rng.integers(0, A, size=n)]
  kind = rng.integers(0, 3)
  if kind == 0:
    rewards = [float(np.sign(x)) if abs(x) > 1.3 else 0.0 for x in rng.normal(size=n)]
  elif kind == 1:
    rewards = [float(x) for x in rng.normal(0, 1, size=n)]
  else:
    rewards = [int(x) for x in rng.integers(-1, 2, size=n)]  # python ints, as gym may return
  errors = [float(x) for x in rng.normal(0, 1, size=n)]
  dones = [False] * n
  steps = list(range(n))
  to_play = [1 if (not two_players or i % 2 == 0) else -1 for i in range(n)]
  return ref_game.HistorySlice(obs, cv, root_values, actions, rewards, errors, dones, steps,
                               [None] * n, to_play)


REPLAY_CASES = {
    "breakout": dict(cfg=dict(action_space=4, td_steps=10, num_unroll_steps=5, discount=0.997,
                              batch_size=64, obs_space=(128,), window_size=2048, beta=0.4),
                     lens=[60, 37, 5, 120, 3, 90, 200], obs_uint8=True, two_players=False),
    "lunar_td1000": dict(cfg=dict(action_space=4, td_steps=1000, num_unroll_steps=5,
                                  discount=0.997, batch_size=32, obs_space=(8,), window_size=4096,
                                  alpha=0.6, beta=1.0),
                         lens=[1300, 40, 700, 1005, 1006, 12], obs_uint8=False, two_players=False),
    "ttt": dict(cfg=dict(action_space=9, td_steps=10, num_unroll_steps=5, discount=1.0,
                         batch_size=48, obs_space=(9,), window_size=256, window_step=64,
                         two_players=True),
                lens=[9, 5, 7, 9, 6, 8, 9, 9, 5, 7, 9, 8, 6, 9, 9, 7, 5, 9, 9, 9, 8, 7, 9, 9, 6, 9,
                      9, 9, 9, 9, 9, 9, 9, 9, 9, 9, 9, 9, 9, 9], obs_uint8=False,
                two_players=True),
}


def import_reference_wrappers():
  """wrappers.py needs gym and cv2 at import time only (base classes and one cv2 switch); stub modules
  let the UNMODIFIED file import so that ClipRewardEnv.reward (wrappers.py:236-238) itself is what
  clips the rewards of the `clip` fixture."""
  gym = types.ModuleType("gym")
  for cls in ("Wrapper", "RewardWrapper", "ObservationWrapper", "ActionWrapper", "Env"):
    setattr(gym, cls, type(cls, (object,), {}))
  gym.spaces = types.ModuleType("gym.spaces")
  cv2 = types.ModuleType("cv2")
  cv2.ocl = types.SimpleNamespace(setUseOpenCL=lambda flag: None)
  sys.modules.setdefault("gym", gym)
  sys.modules.setdefault("gym.spaces", gym.spaces)
  sys.modules.setdefault("cv2", cv2)
  import wrappers as ref_wrappers
  return ref_wrappers


def gen_replay(rng):
  for name, case in REPLAY_CASES.items():
    gen_replay_case(rng, name, case)


def gen_replay_clip():
  """Breakout-shaped replay whose histories carry RAW environment rewards (floats of any size, python
  ints, +0.0 / -0.0); the reference sees them through its own ClipRewardEnv.reward, the product gets the
  raw rewards plus clip_rewards=True.  Own seed: the fixtures above stay bit-identical."""
  ref_wrappers = import_reference_wrappers()
  clip_env = types.SimpleNamespace()
  clip = lambda r: ref_wrappers.ClipRewardEnv.reward(clip_env, r)
  case = dict(cfg=dict(action_space=4, td_steps=10, num_unroll_steps=5, discount=0.997, batch_size=64,
                       obs_space=(128,), window_size=2048, beta=0.4),
              lens=[80, 33, 4, 150, 61], obs_uint8=True, two_players=False)
  gen_replay_case(np.random.default_rng(20261018), "breakout_clip", case, clip=clip)


def gen_replay_case(rng, name, case, clip=None):
  if True:
    cfg = make_config(**case["cfg"])
    A, K, B = cfg.action_space, cfg.num_unroll_steps, cfg.batch_size
    obs_dim = cfg.obs_space[0]
    rb = ref_replay.PrioritizedReplay(cfg)
    hists, ignores, aliases = [], [], []
    overlap = cfg.num_unroll_steps + cfg.td_steps
    for i, n in enumerate(case["lens"]):
      running = (i % 3 == 1) and n > overlap
      h = synth_history(rng, n, A, obs_dim, case["two_players"], case["obs_uint8"], running)
      ignore = overlap if running else None
      if clip is not None:
        raw = [float(x) * 3 for x in rng.normal(size=n)]
        for j in range(0, n, 7):
          raw[j] = [0.0, -0.0, 2, -3, 0, 1e-30, -1e-30][(j // 7) % 7]  # zeros of both signs, python ints, tiny
        seen = h._replace(rewards=[clip(r) for r in raw])                # what the env wrapper hands on
        rb.save_history(seen, ignore=ignore, terminal=not running)
        hist_alias = seen
        h = h._replace(rewards=raw)
      else:
        rb.save_history(h, ignore=ignore, terminal=not running)
        hist_alias = h
      aliases.append(hist_alias)
      hists.append(h)
      ignores.append(-1 if ignore is None else ignore)
    hist_index = {id(h): i for i, h in enumerate(aliases)}

    batches = []
    real_uniform, real_randint = ref_replay.random.uniform, np.random.randint
    try:
      for it in range(3):
        frac = rng.random(B)
        pads = rng.integers(0, A, size=(B, K))
        state = {"b": 0, "k": [0] * B}

        def fake_uniform(s1, s2):
          b = state["b"]
          state["b"] += 1
          return s1 + (s2 - s1) * frac[b]

        def fake_randint(n):
          b = state["b"] - 1
          k = state["k"][b]
          state["k"][b] += 1
          return int(pads[b, k])

        ref_replay.random.uniform = fake_uniform
        np.random.randint = fake_randint
        # which (history, step) each row lands on: replay the same descent through get_leaf
        total = rb.tree.total_priority
        seg = total / B
        picks = []
        for b in range(B):
          v = seg * b + (seg * (b + 1) - seg * b) * frac[b]
          idx, pr, step, h = rb.tree.get_leaf(v)
          picks.append((idx, pr, step, hist_index[id(h)]))
        beta_before = rb.beta
        (obs, actions, (t_r, t_v, t_p)), idxs, is_w = rb.sample_batch()
        # sample_batch mutates history.actions when padding (replay_buffer.py:149-151 appends to a
        # slice copy, so the stored list is untouched) -- nothing to undo.
        assert [p[0] for p in picks] == list(idxs)
        batches.append(dict(
            frac=frac, pads=pads, idxs=np.array(idxs, np.int64),
            priorities=np.array([p[1] for p in picks]), steps=np.array([p[2] for p in picks], np.int32),
            hist=np.array([p[3] for p in picks], np.int32), obs=obs,
            actions=np.array(actions, np.int32), t_rewards=t_r, t_values=t_v, t_policies=t_p,
            is_weights=is_w, beta_before=beta_before, beta_after=rb.beta, total_priority=total,
            num_memories=rb.tree.num_memories))
        # priority update with fresh errors (learners.py:182-184)
        errs = rng.normal(0, 1, size=B).astype(np.float32)
        rb.update(idxs, errs)
        batches[-1]["update_errors"] = errs
        batches[-1]["tree_after_update"] = rb.tree.tree.copy()
    finally:
      ref_replay.random.uniform = real_uniform
      np.random.randint = real_randint

    save = dict(action_space=np.int32(A), num_unroll_steps=np.int32(K), td_steps=np.int32(cfg.td_steps),
                discount=np.float64(cfg.discount), batch_size=np.int32(B), obs_dim=np.int32(obs_dim),
                window_size=np.int32(cfg.window_size),
                window_step=np.int32(cfg.window_size if cfg.window_step is None else cfg.window_step),
                alpha=np.float64(cfg.alpha), beta=np.float64(case["cfg"].get("beta", 1.0)),
                beta_increment=np.float64(cfg.beta_increment_per_sampling),
                epsilon=np.float64(cfg.epsilon), obs_uint8=np.int32(case["obs_uint8"]),
                n_hist=np.int32(len(hists)), ignores=np.array(ignores, np.int32),
                n_batches=np.int32(len(batches)))
    if clip is not None:
      save["clip_rewards"] = np.int32(1)
    for i, h in enumerate(hists):
      save["h%d_obs" % i] = np.stack(h.observations)
      save["h%d_child_visits" % i] = np.array(h.child_visits, np.float64).reshape(-1, A)
      save["h%d_root_values" % i] = np.array(h.root_values, np.float64)
      save["h%d_actions" % i] = np.array(h.actions, np.int32)
      save["h%d_rewards" % i] = np.array(h.rewards, np.float64)
      save["h%d_errors" % i] = np.array(h.errors, np.float64)
      save["h%d_to_play" % i] = np.array(h.to_play, np.int8)
    for i, b in enumerate(batches):
      for k, v in b.items():
        save["b%d_%s" % (i, k)] = np.asarray(v)
    np.savez_compressed(os.path.join(HERE, "replay_%s.npz" % name), **save)
    print("replay", name, "memories", rb.tree.num_memories)


# --------------------------------------------------------------------------------------------
# transforms and FCNetwork
# --------------------------------------------------------------------------------------------
def gen_transforms(rng):
  cfg = make_config()
  x = np.concatenate([
      rng.normal(0, 3, size=400), rng.normal(0, 40, size=200), np.arange(-17, 18, dtype=np.float64),
      np.array([0.0, -0.0, 1e-8, -1e-8, 14.999, 15.0, 15.001, -15.0, 0.5, -0.5, 300.0, -300.0])
  ]).astype(np.float32)
  xt = torch.from_numpy(x.copy())
  h = ref_config.Config.scalar_transform(xt).numpy()
  x2 = np.concatenate([h, rng.uniform(-16, 16, size=300).astype(np.float32)]).reshape(-1, 1)
  sup_in = torch.from_numpy(x2.copy())
  support = cfg.value_phi(sup_in).numpy()  # clamps sup_in in place
  logits = (rng.normal(0, 2, size=(512, 31)) * rng.choice([0.1, 1.0, 4.0], size=(512, 1))).astype(
      np.float32)
  inv = cfg.inverse_value_transform(torch.from_numpy(logits.copy())).numpy()
  cfg_nt = make_config(no_target_transform=True)
  inv_nt = cfg_nt.inverse_value_transform(torch.from_numpy(logits.copy())).numpy()
  # h^-1 alone on exact float32 inputs (the expectation), to pin the float32 op order bit for bit
  v = np.concatenate([rng.uniform(-15, 15, size=2000), [0.0, -0.0, 1.0, -1.0, 15.0, -15.0]]).astype(
      np.float32).reshape(-1, 1)
  onehot_logits = None
  vt = torch.from_numpy(v.copy())
  hinv = (torch.sign(vt) * (((torch.sqrt(1 + 4 * 0.001 * (torch.abs(vt) + 1 + 0.001)) - 1) /
                             (2 * 0.001)) ** 2 - 1)).numpy()  # the expression at config.py:32
  np.savez_compressed(os.path.join(HERE, "transforms.npz"), x=x, h=h, support_in=x2,
                      support=support, logits=logits, inverse=inv, inverse_no_transform=inv_nt,
                      hinv_in=v, hinv_out=hinv)
  print("transforms ok")


def gen_fcnet(rng):
  torch.manual_seed(1234)
  for name, (obs_dim, A, B) in {"atari18": (128, 18, 64), "ttt": (9, 9, 16)}.items():
    cfg = make_config(action_space=A)
    net = ref_networks.FCNetwork(obs_dim, A, torch.device("cpu"), cfg)
    with torch.no_grad():  # make LayerNorm non-trivial
      net.LN.weight.copy_(torch.from_numpy(rng.uniform(0.5, 1.5, size=50).astype(np.float32)))
      net.LN.bias.copy_(torch.from_numpy(rng.normal(0, 0.1, size=50).astype(np.float32)))
    net.eval()
    obs = torch.from_numpy(rng.normal(size=(B, obs_dim)).astype(np.float32))
    actions = [int(a) for a in rng.integers(0, A, size=B)]
    with torch.inference_mode():
      init = net.initial_inference(obs)
      rec = net.recurrent_inference(init.hidden_state, actions)
      net.train()
      rec_train = net.recurrent_inference(init.hidden_state, actions)
      net.eval()
    save = {"w_" + k: v.numpy() for k, v in net.state_dict().items()}
    np.savez_compressed(
        os.path.join(HERE, "fcnet_%s.npz" % name), obs=obs.numpy(), actions=np.array(actions, np.int32),
        init_value=init.value.numpy(), init_logits=init.policy_logits.numpy(),
        init_hidden=init.hidden_state.numpy(), rec_value=rec.value.numpy(),
        rec_reward=rec.reward.numpy(), rec_logits=rec.policy_logits.numpy(),
        rec_hidden=rec.hidden_state.numpy(), rec_value_logits=rec_train.value.numpy(),
        rec_reward_logits=rec_train.reward.numpy(), **save)
    print("fcnet", name, "params", sum(p.numel() for p in net.parameters()))


def gen_fcnet_nosupport():
  """FCNetwork with `--no_support` (config.py:95; networks.py:135-136, 153, 161): one-unit value / reward heads whose
  raw outputs are the scalars.  Own generator so that the other fixtures keep their draws."""
  rng = np.random.default_rng(20261018)
  torch.manual_seed(4321)
  obs_dim, A, B = 8, 4, 48
  cfg = make_config(action_space=A, no_support=True)
  net = ref_networks.FCNetwork(obs_dim, A, torch.device("cpu"), cfg)
  with torch.no_grad():
    net.LN.weight.copy_(torch.from_numpy(rng.uniform(0.5, 1.5, size=50).astype(np.float32)))
    net.LN.bias.copy_(torch.from_numpy(rng.normal(0, 0.1, size=50).astype(np.float32)))
  net.eval()
  obs = torch.from_numpy(rng.normal(size=(B, obs_dim)).astype(np.float32))
  actions = [int(a) for a in rng.integers(0, A, size=B)]
  with torch.inference_mode():
    init = net.initial_inference(obs)
    rec = net.recurrent_inference(init.hidden_state, actions)
  assert init.value.shape == (B, 1) and rec.reward.shape == (B, 1)
  save = {"w_" + k: v.numpy() for k, v in net.state_dict().items()}
  np.savez_compressed(
      os.path.join(HERE, "fcnet_nosupport.npz"), obs=obs.numpy(), actions=np.array(actions, np.int32),
      init_value=init.value.numpy(), init_logits=init.policy_logits.numpy(), init_hidden=init.hidden_state.numpy(),
      rec_value=rec.value.numpy(), rec_reward=rec.reward.numpy(), rec_logits=rec.policy_logits.numpy(),
      rec_hidden=rec.hidden_state.numpy(), **save)
  print("fcnet no_support params", sum(p.numel() for p in net.parameters()))


def gen_muzero():
  """MuZeroNetwork (networks.py:393-554) in eval mode on weights drawn by
  oracle.muzero_ref.seeded_state_dict (23 M parameters are not committed, only the seed)."""
  from oracle import muzero_ref
  C_in, A, seed, B = 4, 18, 20261017, 2
  cfg = make_config(action_space=A)
  net = ref_networks.MuZeroNetwork(C_in, A, "cpu", cfg)
  net.load_state_dict(muzero_ref.seeded_state_dict(C_in, A, seed))
  net.eval()
  g = torch.Generator().manual_seed(seed + 1)
  obs = torch.rand((B, C_in, 96, 96), generator=g)
  actions = [3, 17]
  with torch.inference_mode():
    init = net.initial_inference(obs)
    rec = net.recurrent_inference(init.hidden_state, actions)
    rec2 = net.recurrent_inference(rec.hidden_state, [0, 9])
  np.savez_compressed(
      os.path.join(HERE, "muzero_net.npz"), input_channels=np.int32(C_in), action_space=np.int32(A),
      seed=np.int64(seed), obs=obs.numpy(), actions=np.array(actions, np.int32),
      actions2=np.array([0, 9], np.int32), init_hidden=init.hidden_state.numpy(),
      init_logits=init.policy_logits.numpy(), init_value=init.value.numpy(),
      rec_hidden=rec.hidden_state.numpy(), rec_logits=rec.policy_logits.numpy(),
      rec_value=rec.value.numpy(), rec_reward=rec.reward.numpy(),
      rec2_hidden=rec2.hidden_state.numpy(), rec2_logits=rec2.policy_logits.numpy(),
      rec2_value=rec2.value.numpy(), rec2_reward=rec2.reward.numpy())
  print("muzero: value", init.value.flatten().tolist(), rec.value.flatten().tolist(), "reward",
        rec.reward.flatten().tolist(), "hidden mean", float(rec.hidden_state.mean()))


def gen_muzero_wide():
  """The benchmarked towers: C_in = 32 (stack_obs = 32) and C_in = 64 (with stack_actions, utils.py:28-32),
  batch 8.  The observations are not stored (9-19 MB): the test redraws them from the same seeded CPU generator."""
  from oracle import muzero_ref
  A, B = 18, 8
  for C_in in (32, 64):
    seed = 20261017 + C_in
    cfg = make_config(action_space=A)
    net = ref_networks.MuZeroNetwork(C_in, A, "cpu", cfg)
    net.load_state_dict(muzero_ref.seeded_state_dict(C_in, A, seed))
    net.eval()
    g = torch.Generator().manual_seed(seed + 1)
    obs = torch.rand((B, C_in, 96, 96), generator=g)
    actions, actions2 = [3, 17, 0, 9, 11, 5, 2, 8], [0, 9, 4, 4, 16, 1, 13, 7]
    with torch.inference_mode():
      init = net.initial_inference(obs)
      rec = net.recurrent_inference(init.hidden_state, actions)
      rec2 = net.recurrent_inference(rec.hidden_state, actions2)
    np.savez_compressed(
        os.path.join(HERE, "muzero_net_c%d.npz" % C_in), input_channels=np.int32(C_in), action_space=np.int32(A),
        seed=np.int64(seed), batch=np.int32(B), obs_checksum=np.float64(obs.double().sum().item()),
        actions=np.array(actions, np.int32), actions2=np.array(actions2, np.int32),
        init_hidden=init.hidden_state.numpy(), init_logits=init.policy_logits.numpy(), init_value=init.value.numpy(),
        rec_hidden=rec.hidden_state.numpy(), rec_logits=rec.policy_logits.numpy(),
        rec_value=rec.value.numpy(), rec_reward=rec.reward.numpy(),
        rec2_hidden=rec2.hidden_state.numpy(), rec2_logits=rec2.policy_logits.numpy(),
        rec2_value=rec2.value.numpy(), rec2_reward=rec2.reward.numpy())
    print("muzero wide", C_in, "value range", float(rec.value.abs().max()), float(rec2.value.abs().max()))


def gen_selfplay():
  """Self-play of Tic-Tac-Toe: the move body and the history hand-off of Actor.play_game
  (actors.py:125-176) restated around the reference's Game, TicTacToe, Node, MCTS and
  Config.select_action (actors.py itself needs ray / gym).  The network is the obs-hash stand-in, the
  random draws are recorded.  One fixture row per move; one row per history pushed to the replay."""
  from custom_environments.tic_tac_toe import TicTacToe
  from model_based_rl_b200.testing import ObsHashNetwork
  A, S, K, T, MAXH, G = 9, 12, 2, 2, 3, 8
  cfg = make_config(action_space=A, num_simulations=S, two_players=True, discount=1.0,
                    known_bounds=[-1, 1], num_unroll_steps=K, td_steps=T, obs_space=(9,))
  net = ObsHashNetwork(A, 0.8, 0.5, 2.0, 3)
  engine = ref_mcts.MCTS(cfg)
  rng = np.random.default_rng(777)
  fake = _U()
  real_choice, real_dirichlet = np.random.choice, np.random.dirichlet
  moves, saves, envs = [], [], []
  np.random.choice = fake
  try:
    for gi in range(G):
      env = TicTacToe()
      game = ref_game.Game(env, cfg)
      temperature = [1.0, 0.5, 0.0, 0.25][gi % 4]
      mv = 0
      while not game.terminal:
        root = ref_mcts.Node(0)
        obs = np.float32(game.get_observation(-1))
        init = net.initial_inference(torch.from_numpy(obs).unsqueeze(0))
        legal = env.legal_actions()
        root.expand(init, game.to_play, legal)
        noise = rng.dirichlet([cfg.root_dirichlet_alpha] * len(legal))
        np.random.dirichlet = lambda alpha, noise=noise: noise.copy()
        root.add_exploration_noise(cfg.root_dirichlet_alpha, cfg.root_exploration_fraction)
        np.random.dirichlet = real_dirichlet
        engine.run(root, net)
        error = root.value() - init.value.item()
        game.history.errors.append(error)
        fake.u = float(rng.random())
        action = cfg.select_action(root, temperature)
        game.apply(action)
        game.store_search_statistics(root)
        dense = np.zeros(A)
        dense[:len(legal)] = noise
        moves.append(dict(game=gi, move=mv, obs=obs, legal=sum(1 << int(a) for a in legal), noise=dense,
                          u=fake.u, temperature=temperature, action=action, error=error,
                          root_value=game.history.root_values[-1],
                          child_visits=np.array(game.history.child_visits[-1]),
                          reward=game.history.rewards[-1], done=game.history.dones[-1],
                          to_play=game.history.to_play[-1], next_obs=np.float32(game.history.observations[-1])))
        save_history = (game.history_idx - game.previous_collect_to) == MAXH
        if save_history or game.done or game.terminal:
          overlap = K + T
          if not game.history.dones[game.previous_collect_to - 1]:
            collect_from = max(0, game.previous_collect_to - overlap)
          else:
            collect_from = game.previous_collect_to
          h = game.get_history_sequence(collect_from)
          ignore = overlap if not game.done else None
          saves.append(dict(game=gi, move=mv, collect_from=collect_from, n=len(h.root_values),
                            n_obs=len(h.observations), ignore=-1 if ignore is None else ignore,
                            terminal=game.terminal, actions=list(h.actions), errors=list(h.errors)))
        mv += 1
      envs.append(dict(result=game.info["result"], length=game.step))
  finally:
    np.random.choice = real_choice
    np.random.dirichlet = real_dirichlet
  out = dict(A=np.int32(A), S=np.int32(S), K=np.int32(K), T=np.int32(T), max_history_length=np.int32(MAXH),
             n_games=np.int32(G), hashnet=np.array([0.8, 0.5, 2.0, 3.0]))
  for k in ("game", "move", "legal", "action"):
    out["m_" + k] = np.array([m[k] for m in moves], np.int64)
  for k in ("u", "temperature", "error", "root_value", "reward"):
    out["m_" + k] = np.array([m[k] for m in moves], np.float64)
  out["m_done"] = np.array([m["done"] for m in moves], bool)
  out["m_to_play"] = np.array([m["to_play"] for m in moves], np.int8)
  for k in ("obs", "noise", "child_visits", "next_obs"):
    out["m_" + k] = np.stack([m[k] for m in moves])
  for k in ("game", "move", "collect_from", "n", "n_obs", "ignore"):
    out["s_" + k] = np.array([s_[k] for s_ in saves], np.int64)
  out["s_terminal"] = np.array([s_["terminal"] for s_ in saves], bool)
  out["s_actions"] = np.array([";".join(map(str, s_["actions"])) for s_ in saves])
  out["g_result"] = np.array([e["result"] for e in envs])
  out["g_length"] = np.array([e["length"] for e in envs], np.int32)
  np.savez_compressed(os.path.join(HERE, "selfplay_ttt.npz"), **out)
  print("selfplay: moves", len(moves), "saves", len(saves), "results", [e["result"] for e in envs])


# --------------------------------------------------------------------------------------------
# learner step (learners.py:164-230)
# --------------------------------------------------------------------------------------------
LEARNER_CASES = {
    # name: (obs_dim, A, B, K, optimizer, clip_grad, no_target_transform)
    "breakout": (128, 4, 48, 5, "RMSprop", 0, False),
    "ttt": (9, 9, 16, 3, "AdamW", 5, False),
    "lunar_raw": (8, 4, 24, 5, "SGD", 0, True),
}


def reference_update_weights(net, cfg, optimizer, batch, clip_grad):
  """The body of Learner.update_weights (learners.py:164-230) around the reference's FCNetwork and
  Config; the two loss closures are utils.get_loss_functions' (utils.py:53-56; utils.py needs gym)."""
  def ce(logits, target):
    return (-target * torch.nn.LogSoftmax(dim=1)(logits)).sum(1)

  (observations, actions, targets), idxs, is_weights = batch
  target_rewards, target_values, target_policies = targets
  observations = torch.from_numpy(observations)
  value, reward, policy_logits, hidden_state = net.initial_inference(observations)
  with torch.no_grad():
    target_policies = torch.from_numpy(target_policies)
    target_values = torch.from_numpy(target_values)
    target_rewards = torch.from_numpy(target_rewards)
    is_weights = torch.from_numpy(is_weights)
    init_value = cfg.inverse_value_transform(value) if not cfg.no_support else value  # learners.py:182
    new_errors = (init_value.squeeze() - target_values[:, 0]).cpu().numpy()
    if not cfg.no_target_transform:
      target_values = cfg.scalar_transform(target_values)
      target_rewards = cfg.scalar_transform(target_rewards)
    if not cfg.no_support:  # learners.py:190-192
      target_values = cfg.value_phi(target_values)
      target_rewards = cfg.reward_phi(target_rewards)
  policy_ce = ce
  if cfg.no_support:  # utils.py:61-70: the scalar heads' loss
    ce = {"MSE": torch.nn.MSELoss(reduction='none'), "Huber": torch.nn.SmoothL1Loss(reduction='none')}[cfg.scalar_loss]
  reward_loss = 0
  value_loss = ce(value.squeeze(), target_values[:, 0])
  policy_loss = policy_ce(policy_logits.squeeze(), target_policies[:, 0])
  for i, action in enumerate(zip(*actions), 1):
    value, reward, policy_logits, hidden_state = net.recurrent_inference(hidden_state, action)
    hidden_state.register_hook(lambda grad: grad * 0.5)
    reward_loss += ce(reward.squeeze(), target_rewards[:, i])
    value_loss += ce(value.squeeze(), target_values[:, i])
    policy_loss += policy_ce(policy_logits.squeeze(), target_policies[:, i])
  reward_loss = (is_weights * reward_loss).mean()
  value_loss = (is_weights * value_loss).mean()
  policy_loss = (is_weights * policy_loss).mean()
  full = reward_loss + value_loss + policy_loss
  full.register_hook(lambda grad: grad * (1 / cfg.num_unroll_steps))
  optimizer.zero_grad()
  full.backward()
  grads = {k: p.grad.detach().clone() for k, p in net.named_parameters()}
  if clip_grad:
    torch.nn.utils.clip_grad_norm_(net.parameters(), clip_grad)
  optimizer.step()
  return (reward_loss.item(), value_loss.item(), policy_loss.item()), new_errors, grads


# --no_support learners (config.py:95-96; utils.py:61-70): name: (..., scalar_loss)
LEARNER_NOSUPPORT_CASES = {
    "nosupport_mse": (8, 4, 24, 5, "AdamW", 0, False, "MSE"),
    "nosupport_huber": (9, 9, 16, 3, "RMSprop", 5, True, "Huber"),
}


def gen_learner(cases=None):
  """Two consecutive Learner.update_weights steps per case; optimisers as utils.get_optimizer
  (utils.py:72-83) with the reference's default hyper-parameters (config.py:183-188)."""
  for name, spec in (LEARNER_CASES if cases is None else cases).items():
    obs_dim, A, B, K, opt_name, clip, no_tt = spec[:7]
    scalar_loss = spec[7] if len(spec) > 7 else None
    rng = np.random.default_rng({"breakout": 1, "ttt": 2, "lunar_raw": 3, "nosupport_mse": 4, "nosupport_huber": 5}[name])
    torch.manual_seed(4321)
    cfg = make_config(action_space=A, num_unroll_steps=K, batch_size=B, no_target_transform=no_tt,
                      **({} if scalar_loss is None else dict(no_support=True, scalar_loss=scalar_loss)))
    net = ref_networks.FCNetwork(obs_dim, A, torch.device("cpu"), cfg)
    with torch.no_grad():
      net.LN.weight.copy_(torch.from_numpy(rng.uniform(0.5, 1.5, size=50).astype(np.float32)))
      net.LN.bias.copy_(torch.from_numpy(rng.normal(0, 0.1, size=50).astype(np.float32)))
    net.train()
    lr, mom, wd = 0.0008, 0.9, 1e-4
    if opt_name == "RMSprop":
      opt = torch.optim.RMSprop(net.parameters(), lr=lr, momentum=mom, eps=0.01, weight_decay=wd)
    elif opt_name == "AdamW":
      opt = torch.optim.AdamW(net.parameters(), lr=lr, weight_decay=wd, eps=0.00015)
    else:
      opt = torch.optim.SGD(net.parameters(), lr=lr, momentum=mom, weight_decay=wd)
    save = {"w0_" + k: v.numpy().copy() for k, v in net.state_dict().items()}
    save.update(obs_dim=np.int32(obs_dim), action_space=np.int32(A), batch=np.int32(B), K=np.int32(K),
                optimizer=np.array(opt_name), clip_grad=np.int32(clip), no_target_transform=np.int32(no_tt),
                lr=np.float64(lr), momentum=np.float64(mom), weight_decay=np.float64(wd))
    if scalar_loss is not None:
      save.update(no_support=np.int32(1), scalar_loss=np.array(scalar_loss))
    for step in range(2):
      obs = rng.normal(size=(B, obs_dim)).astype(np.float32)
      actions = [[int(a) for a in rng.integers(0, A, size=K)] for _ in range(B)]
      t_values = (rng.normal(0, 6, size=(B, K + 1)) * (rng.random((B, K + 1)) < 0.9)).astype(np.float32)
      t_values[0, 0], t_values[1, 1], t_values[2, 2] = 40.0, -3.0, 2.0  # clamp + integers
      t_rewards = np.sign(rng.normal(size=(B, K + 1)) * (rng.random((B, K + 1)) < 0.3)).astype(np.float32)
      t_rewards[3, 1] = 0.37
      pol = rng.random((B, K + 1, A)) ** 3
      pol /= pol.sum(-1, keepdims=True)
      past_end = rng.random((B, K + 1)) < 0.15  # positions past the end of the episode: zeros
      pol[past_end] = 0.0
      t_policies = pol.astype(np.float32)
      is_w = rng.uniform(0.05, 1.0, size=B)
      is_w /= is_w.max()
      batch = ((obs, actions, (t_rewards.copy(), t_values.copy(), t_policies.copy())), None, is_w)
      losses, errs, grads = reference_update_weights(net, cfg, opt, batch, clip)
      save.update({"s%d_obs" % step: obs, "s%d_actions" % step: np.array(actions, np.int32),
                   "s%d_t_values" % step: t_values, "s%d_t_rewards" % step: t_rewards,
                   "s%d_t_policies" % step: t_policies, "s%d_is_weights" % step: is_w,
                   "s%d_losses" % step: np.array(losses, np.float64), "s%d_new_errors" % step: errs})
      for k, g in grads.items():
        g = g.numpy()
        save["s%d_gnorm_%s" % (step, k)] = np.float64(np.sqrt((g.astype(np.float64) ** 2).sum()))
        if g.size <= 2048:
          save["s%d_g_%s" % (step, k)] = g
        else:
          save["s%d_g_%s" % (step, k)] = g.reshape(-1)[::max(1, g.size // 1024)].copy()
      for k, v in net.state_dict().items():
        v = v.numpy()
        save["s%d_w_%s" % (step, k)] = v.copy() if v.size <= 2048 else v.reshape(-1)[::max(1, v.size // 1024)].copy()
    np.savez_compressed(os.path.join(HERE, "learner_%s.npz" % name), **save)
    print("learner", name, "losses", losses)


if __name__ == "__main__":
  torch.set_num_threads(1)
  if len(sys.argv) > 1 and sys.argv[1] == "muzero":
    gen_muzero()
    sys.exit(0)
  if len(sys.argv) > 1 and sys.argv[1] == "selfplay":
    gen_selfplay()
    sys.exit(0)
  if len(sys.argv) > 1 and sys.argv[1] == "muzero_wide":
    gen_muzero_wide()
    sys.exit(0)
  if len(sys.argv) > 1 and sys.argv[1] == "clip":
    gen_replay_clip()
    sys.exit(0)
  if len(sys.argv) > 1 and sys.argv[1] == "fcnet_nosupport":
    gen_fcnet_nosupport()
    sys.exit(0)
  if len(sys.argv) > 1 and sys.argv[1] == "learner":
    gen_learner()
    sys.exit(0)
  if len(sys.argv) > 1 and sys.argv[1] == "learner_nosupport":
    gen_learner(LEARNER_NOSUPPORT_CASES)
    sys.exit(0)
  rng = np.random.default_rng(20261017)
  gen_search(rng)
  gen_select_action(rng)
  gen_replay(rng)
  gen_transforms(rng)
  gen_fcnet(rng)
  gen_muzero()
  gen_selfplay()
  gen_learner()
  gen_replay_clip()
  gen_muzero_wide()
  gen_fcnet_nosupport()
  gen_learner(LEARNER_NOSUPPORT_CASES)
  print("python", sys.version.split()[0], "numpy", np.__version__, "torch", torch.__version__)
