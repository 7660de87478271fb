"""Pure-Python port of the reference's per-move search path -- TEST INFRASTRUCTURE / CPU BASELINE.

This is the `"port"` that bench.py times as `cpu_baseline` and as `--impl reference`: the same
work the reference's Ray actor does for one move (actors.py:131-153), in the same language with
the same arithmetic (python floats, math.exp / math.log / math.sqrt, one batch-size-1
recurrent_inference per simulation), restated over flat per-game arrays instead of Node objects.
It is pinned against the unmodified reference by tests/test_oracle_golden.py (golden vectors) and
is never imported by the product path.

Reference: mcts.py:6-145, config.py:70-81, game.py:106-115, actors.py:131-153.
"""
import math

import numpy as np


class FlatSearch(object):
  """MCTS.run for one game; children of node n live at [n * A, (n + 1) * A) in flat lists."""

  def __init__(self, num_simulations, action_space, two_players=False, discount=0.997,
               pb_c_base=19652, pb_c_init=1.25, init_value_score=0.0, known_bounds=(None, None)):
    self.S, self.A = int(num_simulations), int(action_space)
    self.two_players = bool(two_players)
    self.discount = discount
    self.pb_c_base, self.pb_c_init = pb_c_base, pb_c_init
    self.init_value_score = init_value_score
    self.known_bounds = tuple(known_bounds)

  # -- Node.expand (mcts.py:47-55) -------------------------------------------------------------------
  def _expand(self, n, logits_row, actions):
    A = self.A
    w = {a: math.exp(logits_row[a].item()) for a in actions}
    total = sum(w.values())
    base = n * A
    for a, x in w.items():
      self.prior[base + a] = x / total
      self.child[base + a] = -1

  def setup_root(self, init_out, to_play, legal_actions, noise=None, frac=0.25):
    """root.expand + add_exploration_noise (actors.py:142-143, mcts.py:57-61)."""
    S, A = self.S, self.A
    n_edges = (S + 1) * A
    self.prior = [0.0] * n_edges
    self.child = [-2] * n_edges          # -2: no such child, -1: unexpanded, >= 0: node id
    self.visit = [0] * (S + 1)           # per node (root = 0)
    self.vsum = [0] * (S + 1)
    self.reward = [0] * (S + 1)
    self.hidden = [None] * (S + 1)
    self.root_to_play = to_play
    self.hidden[0] = init_out.hidden_state
    self._expand(0, init_out.policy_logits[0], legal_actions)
    if noise is not None:
      for a, z in zip(legal_actions, noise):
        self.prior[a] = self.prior[a] * (1 - frac) + z * frac
    self.trace = []

  # -- MCTS.run (mcts.py:78-102) -----------------------------------------------------------------------
  def run(self, network):
    A, two, disc = self.A, self.two_players, self.discount
    lo = float('inf') if self.known_bounds[0] is None else self.known_bounds[0]
    hi = -float('inf') if self.known_bounds[1] is None else self.known_bounds[1]
    prior, child, visit, vsum, reward = self.prior, self.child, self.visit, self.vsum, self.reward
    for sim in range(self.S):
      node, to_play = 0, self.root_to_play
      path = [0]
      while True:
        base = node * A
        N = visit[node]
        best_key, best_a = None, -1
        for a in range(A):
          c = child[base + a]
          if c == -2:
            continue
          if N == 0:
            key = prior[base + a]
          else:  # ucb_score mcts.py:115-124
            n = visit[c] if c >= 0 else 0
            pb_c = math.log((N + self.pb_c_base + 1) / self.pb_c_base) + self.pb_c_init
            pb_c *= math.sqrt(N) / (n + 1)
            key = pb_c * prior[base + a]
            if n > 0:
              v = vsum[c] / n
              if two:
                v = -v
              q = reward[c] + disc * v
              if hi > lo:
                q = (q - lo) / (hi - lo)
              elif hi == lo:
                q = 1.0
              key += q
            else:
              key += self.init_value_score
          if best_key is None or key >= best_key:   # ties -> larger action (tuple max)
            best_key, best_a = key, a
        if two:
          to_play *= -1
        nxt = child[base + best_a]
        if nxt < 0:
          break
        node = nxt
        path.append(node)
      new = sim + 1
      out = network.recurrent_inference(self.hidden[node], [best_a])
      child[base + best_a] = new
      self.hidden[new] = out.hidden_state
      if out.reward:
        reward[new] = out.reward.item()
      self._expand(new, out.policy_logits[0], range(A))
      path.append(new)
      self.trace.append((node, best_a, len(path) - 1))
      # backpropagate mcts.py:126-143
      value = out.value.item()
      depth = len(path) - 1
      for k in range(depth, -1, -1):
        n = path[k]
        if two:
          node_tp = -self.root_to_play if k % 2 else self.root_to_play
        else:
          node_tp = self.root_to_play
        same = node_tp == to_play
        vsum[n] += value if same else -value
        visit[n] += 1
        r = -reward[n] if (two and same) else reward[n]
        if k > 0:
          mean = vsum[n] / visit[n]
          q = reward[n] - disc * mean if two else reward[n] + disc * mean
          if q < lo:
            lo = q
          if q > hi:
            hi = q
        value = r + disc * value
    self.minmax = (lo, hi)

  # -- what callers read afterwards -----------------------------------------------------------------
  def root_visits(self):
    out = [0] * self.A
    for a in range(self.A):
      c = self.child[a]
      if c >= 0:
        out[a] = self.visit[c]
    return out

  def root_value(self):
    return 0 if self.visit[0] == 0 else self.vsum[0] / self.visit[0]

  def child_visits(self):  # game.py:107-110
    v = self.root_visits()
    s = sum(v[a] for a in range(self.A) if self.child[a] != -2)
    return [v[a] / s if self.child[a] != -2 else 0 for a in range(self.A)]


def select_action(visits, legal_actions, temperature, u):
  """Config.select_action (config.py:70-81) with one host-supplied uniform."""
  counts = np.array([visits[a] for a in legal_actions])
  if temperature:
    d = counts ** (1 / temperature)
    d = d / d.sum()
    cdf = d.cumsum()
    cdf /= cdf[-1]
    idx = int(cdf.searchsorted(u, side='right'))
  else:
    ties = np.where(counts == counts.max())[0]
    idx = int(ties[int(math.floor(u * len(ties)))])
  return legal_actions[idx]


def play_move(search, network, obs_row, noise_row, temperature, u, to_play=1, legal_actions=None,
              frac=0.25):
  """One move of Actor.play_game (actors.py:131-153) for one game; returns (action, root_value,
  child_visits, initial value)."""
  import torch
  legal_actions = list(range(search.A)) if legal_actions is None else legal_actions
  init = network.initial_inference(torch.as_tensor(obs_row).unsqueeze(0))
  search.setup_root(init, to_play, legal_actions, noise_row, frac)
  search.run(network)
  action = select_action(search.root_visits(), legal_actions, temperature, u)
  return action, search.root_value(), search.child_visits(), init.value.item()
