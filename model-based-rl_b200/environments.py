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
