"""ORACLE (test infrastructure, never on the product path): CPU restatement of the reference's
SumTree and of the sampling half of PrioritizedReplay (replay_buffer.py:6-66, 110-145, 160-163,
200-203).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline may import this.

Pinned bit for bit to tests/golden/replay_*.npz (tree contents after every update, sampled tree
indices, priorities, importance weights, num_memories) by tests/test_oracle_golden.py."""
import numpy as np


class SumTreeRef(object):
  """replay_buffer.py:6-66 with the (step, history) payload kept as (history id, step)."""

  def __init__(self, max_capacity, capacity_step):
    self.tree = np.zeros(2 * max_capacity - 1)
    self.slot_hist = np.full(max_capacity, -1, np.int64)
    self.slot_step = np.zeros(max_capacity, np.int64)
    self.max_capacity = max_capacity
    self.capacity_step = capacity_step
    self.capacity = capacity_step
    self.prev_capacity = 0
    self.num_memories = 0
    self.position = 0

  def add(self, priorities, hist_id):  # replay_buffer.py:19-33
    for step, priority in enumerate(priorities):
      idx = self.position + self.max_capacity - 1
      self.slot_hist[self.position] = hist_id
      self.slot_step[self.position] = step
      self.update(idx, priority)
      if self.position >= self.prev_capacity:
        self.num_memories += 1
      self.position = (self.position + 1) % self.capacity
      if self.position == 0:
        self.prev_capacity = self.capacity
        self.capacity = min(self.max_capacity, self.capacity + self.capacity_step)

  def update(self, idx, priority):  # replay_buffer.py:35-41
    change = priority - self.tree[idx]
    self.tree[idx] = priority
    while idx != 0:
      idx = (idx - 1) // 2
      self.tree[idx] += change

  def get_leaf(self, value):  # replay_buffer.py:43-62
    parent = 0
    while True:
      left = 2 * parent + 1
      if left >= len(self.tree):
        break
      if value <= self.tree[left]:
        parent = left
      else:
        value -= self.tree[left]
        parent = left + 1
    slot = parent - self.max_capacity + 1
    return parent, self.tree[parent], int(self.slot_step[slot]), int(self.slot_hist[slot])

  @property
  def total_priority(self):
    return self.tree[0]


def get_priorities(errors, epsilon, alpha):  # replay_buffer.py:110-111
  return np.power((np.abs(errors) + epsilon), alpha)


def sample_indices(tree, batch_size, u01, beta):
  """replay_buffer.py:134-145 + 160-162 with random.uniform(s1, s2) = s1 + (s2 - s1) * u."""
  seg = tree.total_priority / batch_size
  picks = []
  for b in range(batch_size):
    s1, s2 = seg * b, seg * (b + 1)
    picks.append(tree.get_leaf(s1 + (s2 - s1) * u01[b]))
  priorities = np.array([p[1] for p in picks])
  probs = priorities / tree.total_priority
  w = np.power(tree.num_memories * probs, -beta)
  w /= w.max()
  return picks, w
