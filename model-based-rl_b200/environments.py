"""Vectorised environments for the batched self-play driver.

`VectorTicTacToe` plays G independent games of the reference's custom_environments/tic_tac_toe.py in
lock step on the host (numpy): same board encoding, same observation (`turn * board` after the
move), same reward / terminal / result rules.  Environments are host code in the reference too; they
are not on the GPU hot path.
"""
import numpy as np

# the eight lines of the board (rows, columns, diagonals), as cell indices
_LINES = np.array([[0, 1, 2], [3, 4, 5], [6, 7, 8], [0, 3, 6], [1, 4, 7], [2, 5, 8], [0, 4, 8], [2, 4, 6]])
# _THROUGH[cell] = which lines contain that cell (did_win only inspects lines through the move,
# tic_tac_toe.py:52-70)
_THROUGH = np.zeros((9, 8), dtype=bool)
for _l, _cells in enumerate(_LINES):
  _THROUGH[_cells, _l] = True

RESULTS = ("player 1 wins", "player 2 wins", "draw")


class VectorTicTacToe(object):

  num_actions = 9
  obs_shape = (9,)
  two_players = True

  def __init__(self, num_games):
    self.num_games = int(num_games)
    self.board = np.zeros((self.num_games, 9), dtype=np.int32)
    self.elapsed = np.zeros(self.num_games, dtype=np.int32)
    self.turn = np.ones(self.num_games, dtype=np.int32)

  def reset(self, which=None):
    """tic_tac_toe.py:46-50 for the selected games (all by default); returns their observations."""
    idx = np.arange(self.num_games) if which is None else np.asarray(which)
    self.board[idx] = 0
    self.elapsed[idx] = 0
    self.turn[idx] = 1
    return self.board[idx].copy()

  def legal_mask(self):
    """bit a set <=> cell a is empty (tic_tac_toe.py:41-42)."""
    return ((self.board == 0) << np.arange(9)).sum(axis=1).astype(np.uint32)

  def legal_actions(self, g):
    return np.where(self.board[g] == 0)[0]

  def step(self, actions):
    """tic_tac_toe.py:17-39 for every game: -> (obs [G, 9] int32, reward [G], done [G] bool,
    result [G] index into RESULTS or -1)."""
    actions = np.asarray(actions, dtype=np.int64)
    g = np.arange(self.num_games)
    if (self.board[g, actions] != 0).any():
      raise ValueError("illegal move: the cell is taken")
    self.board[g, actions] = self.turn
    sums = np.abs(self.board[:, _LINES].sum(axis=2))            # [G, 8]
    won = ((sums == 3) & _THROUGH[actions]).any(axis=1)
    draw = ~won & (self.elapsed == 8)
    reward = won.astype(np.int32)
    result = np.where(won, np.where(self.turn == 1, 0, 1), np.where(draw, 2, -1))
    self.elapsed += 1
    self.turn *= -1
    obs = self.turn[:, None] * self.board
    return obs.copy(), reward, won | draw, result


class SyntheticFrames(object):
  """G independent synthetic single-player episodes with image observations [C, H, W] float32 in [0, 1)
  (the shape BASELINE.json's conv configuration names): every action is legal, rewards are seeded noise
  in {-1, 0, 1}, an episode ends after `episode_length` steps.  Stands in for the Atari wrappers
  (wrappers.py), which are environment code and not on the GPU hot path."""

  two_players = False

  def __init__(self, num_games, num_actions, obs_shape, episode_length=8, seed=0):
    self.num_games, self.num_actions, self.obs_shape = int(num_games), int(num_actions), tuple(obs_shape)
    self.episode_length = int(episode_length)
    self.rng = np.random.default_rng(seed)
    self.elapsed = np.zeros(self.num_games, dtype=np.int32)

  def _frames(self, n):
    return self.rng.random((n,) + self.obs_shape, dtype=np.float32)

  def reset(self, which=None):
    idx = np.arange(self.num_games) if which is None else np.asarray(which)
    self.elapsed[idx] = 0
    return self._frames(len(idx))

  def legal_mask(self):
    return np.full(self.num_games, (1 << self.num_actions) - 1, dtype=np.uint32)

  def step(self, actions):
    actions = np.asarray(actions)
    if ((actions < 0) | (actions >= self.num_actions)).any():
      raise ValueError("action outside the action space")
    self.elapsed += 1
    reward = self.rng.integers(-1, 2, size=self.num_games).astype(np.int32)
    done = self.elapsed >= self.episode_length
    return self._frames(self.num_games), reward, done, np.full(self.num_games, -1)


class SyntheticRam(object):
  """G independent synthetic single-player episodes with byte observations [obs_dim] (the Breakout-ram /
  Atari-sweep shape of BASELINE.json: 128 bytes, README.md:56): every action is legal, rewards are seeded draws
  from {-1, 0, 1} (P(nonzero) = 0.1), an episode ends after `episode_length` steps.  Stands in for the ALE
  wrappers (wrappers.py), which are environment code and not on the GPU hot path."""

  two_players = False

  def __init__(self, num_games, num_actions, obs_dim=128, episode_length=600, seed=0):
    self.num_games, self.num_actions, self.obs_shape = int(num_games), int(num_actions), (int(obs_dim),)
    self.episode_length = int(episode_length)
    self.rng = np.random.default_rng(seed)
    self.elapsed = np.zeros(self.num_games, dtype=np.int32)
    self._all = np.full(self.num_games, (1 << self.num_actions) - 1 if self.num_actions < 32 else 0xffffffff,
                        dtype=np.uint32)
    # the frames of a step come from a small seeded pool (drawing 128 bytes for each of 4096 games costs 0.4 ms,
    # as much as a third of the search it feeds; an emulator's step is not what this stand-in measures)
    self._pool = self.rng.integers(0, 256, size=(16, self.num_games) + self.obs_shape, dtype=np.uint8)
    self._t = 0

  def _frames(self, n):
    if n == self.num_games:
      self._t += 1
      return self._pool[self._t % len(self._pool)].copy()
    return self.rng.integers(0, 256, size=(n,) + self.obs_shape, dtype=np.uint8)

  def reset(self, which=None):
    idx = np.arange(self.num_games) if which is None else np.asarray(which)
    self.elapsed[idx] = 0
    return self._frames(len(idx))

  def legal_mask(self):
    return self._all

  def step(self, actions):
    self.elapsed += 1
    reward = (self.rng.integers(-1, 2, size=self.num_games) * (self.rng.random(self.num_games) < 0.1)).astype(np.int32)
    done = self.elapsed >= self.episode_length
    return self._frames(self.num_games), reward, done, np.full(self.num_games, -1)
