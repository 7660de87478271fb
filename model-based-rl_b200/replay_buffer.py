"""Prioritized replay with a device-resident window and fused target construction.

Mirrors the reference's `PrioritizedReplay` (replay_buffer.py:69-210): `save_history(history,
ignore, terminal)`, `sample_batch()` -> `((observations, actions, (target_rewards, target_values,
target_policies)), idxs, is_weights)`, `update(idxs, errors)`, `size()`, `get_throughput()`,
`add_initial_throughput()`; the random draws come from the same generators in the same order
(`random.uniform` per row, `np.random.randint` per padded action), so with equal seeds and equal
contents it samples the same rows as the reference.

What is different underneath: trajectories live in HBM as a structure of arrays over window
positions instead of pickled Python lists, the per-row Python loop of
`sample_batch` + `insert_target` (replay_buffer.py:138-158, 165-198) is one launch of
`mz_build_targets`, and the sum-tree lives in HBM too: stratified `get_leaf`, importance weights and
batched priority updates are CUDA kernels (`mz_sumtree_*`) whose float64 sums stay bit-identical
to the reference's one-leaf-at-a-time loop.
`sample_batch_device()` returns CUDA tensors (optionally with h(x) + two-hot supports fused,
learners.py:186-192) for a learner that consumes them in place.
"""
import bisect
import ctypes as C
import random
from collections import deque

import numpy as np
import torch

from . import _lib


class RingCursor(object):
  """The slot ring of the reference's SumTree.add (replay_buffer.py:19-33): which slot the next
  memory goes to, the growing capacity (`window_step`) and `num_memories`.  Pure host integers."""

  def __init__(self, max_capacity, capacity_step):
    self.max_capacity = int(max_capacity)
    self.capacity_step = int(capacity_step)
    self.capacity = self.capacity_step
    self.prev_capacity = 0
    self.num_memories = 0
    self.position = 0

  def take(self, n):
    """Slots of the next n memories, in order (one arange per stretch up to the next wrap of the ring)."""
    slots = np.empty(n, np.int64)
    done = 0
    while done < n:
      run = min(n - done, self.capacity - self.position)
      slots[done:done + run] = np.arange(self.position, self.position + run)
      # slots at or beyond the previous capacity are new memories (replay_buffer.py:27-28)
      self.num_memories += max(0, self.position + run - max(self.position, self.prev_capacity))
      self.position += run
      done += run
      if self.position == self.capacity:  # wrapped: the ring grows by one step (replay_buffer.py:30-33)
        self.position = 0
        self.prev_capacity = self.capacity
        self.capacity = min(self.max_capacity, self.capacity + self.capacity_step)
    return slots


class ReplayIndex(object):
  """Device-resident SumTree: float64 sums in HBM plus, per slot, where its (history, step) lives
  in the replay window.  The host keeps only the ring cursor and which chunk owns each slot."""

  def __init__(self, max_capacity, capacity_step, device):
    self.lib = _lib.load()
    self.device = device
    self.max_capacity = int(max_capacity)
    self.ring = RingCursor(max_capacity, capacity_step)
    self.tree = torch.zeros(2 * self.max_capacity - 1, dtype=torch.float64, device=device)
    self.slot_pos = torch.zeros(self.max_capacity, dtype=torch.int64, device=device)
    self.slot_start = torch.zeros(self.max_capacity, dtype=torch.int64, device=device)
    self.slot_len = torch.zeros(self.max_capacity, dtype=torch.int32, device=device)
    self.slot_chunk = np.full(self.max_capacity, -1, np.int64)  # host: liveness of window chunks

  @property
  def num_memories(self):
    return self.ring.num_memories

  @property
  def total_priority(self):
    return float(self.tree[0].item())

  def _stage(self, tree_idx, priorities):
    idx = torch.from_numpy(np.ascontiguousarray(tree_idx, np.int64)).to(self.device)
    pri = torch.from_numpy(np.ascontiguousarray(priorities, np.float64)).to(self.device)
    return idx, pri, torch.empty_like(pri)

  def add(self, priorities, chunk_id, chunk_start, chunk_len):
    """SumTree.add (replay_buffer.py:19-33); returns {chunk id: how many of its slots were overwritten} (this
    chunk's own id included when one add laps the ring).

    A history with more memories than the ring currently holds wraps it, so the same slot appears more
    than once in the add.  The reference writes one memory at a time -- the last write to a slot wins and
    every earlier one is an ordinary overwrite -- so the add is applied in pieces without repeated slots."""
    n = len(priorities)
    if n == 0:
      return {}
    slots = self.ring.take(n)
    pri_all = np.ascontiguousarray(priorities, np.float64)
    overwritten = {}
    lo = 0
    while lo < n:
      # longest run from `lo` without a repeated slot
      seg = slots[lo:]
      _, first = np.unique(seg, return_index=True)
      repeat = np.ones(seg.size, bool)
      repeat[first] = False
      rep = np.nonzero(repeat)[0]
      hi = lo + (int(rep[0]) if rep.size else seg.size)
      piece = slots[lo:hi]
      old = self.slot_chunk[piece]
      ids, counts = np.unique(old[old >= 0], return_counts=True)
      for c, k in zip(ids.tolist(), counts.tolist()):
        overwritten[c] = overwritten.get(c, 0) + k
      self.slot_chunk[piece] = chunk_id
      idx, pri, scratch = self._stage(piece + self.max_capacity - 1, pri_all[lo:hi])
      _lib.check(self.lib.mz_sumtree_add_from(_lib.ptr(self.tree), self.max_capacity, hi - lo, _lib.ptr(idx),
                                              _lib.ptr(pri), int(chunk_start), int(chunk_len), lo,
                                              _lib.ptr(self.slot_pos), _lib.ptr(self.slot_start),
                                              _lib.ptr(self.slot_len), _lib.ptr(scratch),
                                              _lib.current_stream()), "mz_sumtree_add_from")
      lo = hi
    return overwritten

  def add_chunks(self, items):
    """SumTree.add for the memories of several chunks, in order, with one upload and one call: items =
    [(priorities, chunk id, chunk start, chunk length)].  Returns {chunk id: overwritten slots} like `add`."""
    items = [it for it in items if len(it[0])]
    total = sum(len(it[0]) for it in items)
    overwritten = {}
    if total == 0:
      return overwritten
    if len(items) == 1 or total > self.ring.capacity:  # an add that laps the ring goes in pieces (see add)
      for pri, cid, start, n in items:
        for c, k in self.add(pri, cid, start, n).items():
          overwritten[c] = overwritten.get(c, 0) + k
      return overwritten
    slots = self.ring.take(total)
    old = self.slot_chunk[slots]
    ids, counts = np.unique(old[old >= 0], return_counts=True)
    overwritten = dict(zip(ids.tolist(), counts.tolist()))
    ns = len(items)
    # one blob: tree idx [total] i64 | priority [total] f64 | chunk start [ns] i64 | begin [ns + 1] i32 | length [ns] i32
    blob = np.empty(16 * total + 8 * ns + 4 * (2 * ns + 2), np.uint8)
    h_idx = blob[:8 * total].view(np.int64)
    h_pri = blob[8 * total:16 * total].view(np.float64)
    h_start = blob[16 * total:16 * total + 8 * ns].view(np.int64)
    h_begin = blob[16 * total + 8 * ns:16 * total + 8 * ns + 4 * (ns + 1)].view(np.int32)
    h_len = blob[16 * total + 8 * ns + 4 * (ns + 1):16 * total + 8 * ns + 4 * (2 * ns + 1)].view(np.int32)
    np.add(slots, self.max_capacity - 1, out=h_idx)
    off = 0
    for s, (pri, cid, start, n) in enumerate(items):
      k = len(pri)
      h_pri[off:off + k] = pri
      self.slot_chunk[slots[off:off + k]] = cid
      h_start[s], h_begin[s], h_len[s] = start, off, n
      off += k
    h_begin[ns] = total
    d = torch.from_numpy(blob).to(self.device)
    base = d.data_ptr()
    scratch = torch.empty(total, dtype=torch.float64, device=self.device)
    P = C.c_void_p
    _lib.check(self.lib.mz_sumtree_add_chunks(_lib.ptr(self.tree), self.max_capacity, total, P(base), P(base + 8 * total),
                                              ns, P(base + 16 * total + 8 * ns), P(base + 16 * total),
                                              P(base + 16 * total + 8 * ns + 4 * (ns + 1)), _lib.ptr(self.slot_pos),
                                              _lib.ptr(self.slot_start), _lib.ptr(self.slot_len), _lib.ptr(scratch),
                                              _lib.current_stream()), "mz_sumtree_add_chunks")
    return overwritten

  @_lib.on_device
  def update(self, tree_idx, priorities):
    """SumTree.update for a batch, in order (replay_buffer.py:200-203)."""
    if len(tree_idx) == 0:
      return
    idx, pri, scratch = self._stage(tree_idx, priorities)
    _lib.check(self.lib.mz_sumtree_update(_lib.ptr(self.tree), self.max_capacity, len(tree_idx),
                                          _lib.ptr(idx), _lib.ptr(pri), _lib.ptr(scratch),
                                          _lib.current_stream()), "mz_sumtree_update")

  def sample(self, u01, beta, with_weights):
    """Stratified get_leaf for len(u01) rows -> device tensors (tree idx, priority, window
    position, chunk start, chunk length, is_weights or None)."""
    n = len(u01)
    dev = self.device
    d_u = torch.from_numpy(np.ascontiguousarray(u01, np.float64)).to(dev)
    blob = torch.empty(6 * n, dtype=torch.int64, device=dev)  # one allocation for the six outputs
    idx, pos, cstart = blob[0:n], blob[n:2 * n], blob[2 * n:3 * n]
    pri = blob[3 * n:4 * n].view(torch.float64)
    isw = blob[4 * n:5 * n].view(torch.float64) if with_weights else None
    clen = blob[5 * n:6 * n].view(torch.int32)[:n]
    self.last_blob = blob  # sample_batch() brings all six outputs to the host with one copy
    _lib.check(self.lib.mz_sumtree_sample(_lib.ptr(self.tree), self.max_capacity, n, _lib.ptr(d_u),
                                          _lib.ptr(self.slot_pos), _lib.ptr(self.slot_start),
                                          _lib.ptr(self.slot_len), self.ring.num_memories,
                                          float(beta), _lib.ptr(idx), _lib.ptr(pri), _lib.ptr(pos),
                                          _lib.ptr(cstart), _lib.ptr(clen), _lib.ptr(isw),
                                          _lib.current_stream()), "mz_sumtree_sample")
    return idx, pri, pos, cstart, clen, isw


class PrioritizedReplay(object):

  def __init__(self, config, device=None, window_positions=None):
    _lib.require_cuda()
    self.lib = _lib.load()
    self.device = _lib.normalize_device(device)
    self.batch_size = int(config.batch_size)
    self.beta_increment_per_sampling = config.beta_increment_per_sampling
    self.epsilon = config.epsilon
    self.alpha = config.alpha
    self.beta = config.beta
    self.num_unroll_steps = int(config.num_unroll_steps)
    self.td_steps = int(config.td_steps)
    self.discount = config.discount
    n_steps = self.num_unroll_steps + self.td_steps
    self.discounts = np.array([self.discount**n for n in range(n_steps)], dtype=np.float32)
    self.action_space = int(config.action_space)
    self.obs_space = tuple(config.obs_space)
    self.obs_elems = int(np.prod(self.obs_space))
    self.target_length = self.num_unroll_steps + 1
    self.value_support = tuple(getattr(config, 'value_support', (-15, 15)))
    self.reward_support = tuple(getattr(config, 'reward_support', (-15, 15)))
    self.no_target_transform = bool(getattr(config, 'no_target_transform', False))
    # config.clip_rewards (config.py:111): the reference clips in an env wrapper (ClipRewardEnv.reward =
    # np.sign, wrappers.py:236-238); here the window keeps what the environment returned and the target
    # kernel reads every reward through sign() -- same targets, and the raw rewards stay available
    self.clip_rewards = bool(getattr(config, 'clip_rewards', False))

    capacity = int(config.window_size)
    step = capacity if getattr(config, 'window_step', None) is None else int(config.window_step)
    self.index = ReplayIndex(capacity, step, self.device)
    self.throughput = {'frames': 0, 'games': 0}
    if getattr(config, 'seed', None) is not None:
      np.random.seed(config.seed)
      random.seed(config.seed + 1)

    # window arena: chunks are appended in a ring; a chunk stays until all its slots are overwritten
    overlap = self.num_unroll_steps + self.td_steps
    chunk_new = int(getattr(config, 'max_history_length', 500))
    if window_positions is None:
      window_positions = int(capacity * (1.0 + overlap / max(1, chunk_new)) * 1.25) + 4 * (chunk_new + overlap)
    self.P = int(window_positions)
    self._obs_dtype = None
    self.w_obs = None
    dev = self.device
    self.w_actions = torch.zeros(self.P, dtype=torch.int32, device=dev)
    self.w_rewards = torch.zeros(self.P, dtype=torch.float32, device=dev)
    self.w_to_play = torch.ones(self.P, dtype=torch.int8, device=dev)
    self.w_root_values = torch.zeros(self.P, dtype=torch.float64, device=dev)
    self.w_child_visits = torch.zeros((self.P, self.action_space), dtype=torch.float32, device=dev)
    self.d_discounts = torch.from_numpy(self.discounts).to(dev)
    self._head = 0
    self.max_chunk = 1          # longest reservation so far (bounds the backwards scan of _overlapping)
    self._chunk_at = {}         # chunk id -> [start, length]
    self._starts = []           # sorted (start, chunk id): which chunks a new allocation would overwrite
    self._live = {}             # chunk id -> referencing slots
    self._next_chunk = 0

  # -- reference API -------------------------------------------------------------------------------
  def add_initial_throughput(self, frames, games):
    self.throughput['frames'] += frames
    self.throughput['games'] += games

  def get_priorities(self, errors):
    return np.power((np.abs(errors) + self.epsilon), self.alpha)

  @_lib.on_device
  def save_history(self, history, ignore=None, terminal=False):
    """replay_buffer.py:113-122.  `history` has the HistorySlice fields (game.py:5-16)."""
    if ignore is not None:
      errors = history.errors[:-ignore]
      priorities = self.get_priorities(errors) if errors else []
    else:
      priorities = self.get_priorities(history.errors)
    n = len(history.root_values)
    if n:
      start = self._upload(history, n)
      cid = self._next_chunk
      self._next_chunk += 1
      self._register(cid, start, n)
      # liveness = sum-tree slots that refer to the chunk: every memory of this history takes one, every
      # overwritten slot gives one back (to an older chunk, or to this one when the add laps the ring)
      self._live[cid] = len(priorities)
      for old, k in self.index.add(priorities, cid, start, n).items():
        if old in self._live:
          self._live[old] -= k
    self.throughput['frames'] += len(priorities)
    if terminal:
      self.throughput['games'] += 1

  # -- device-resident ingest: the self-play driver writes trajectories straight into the window ------
  def _window_struct(self):
    return _lib.Window(self.action_space, self.obs_elems, int(self._obs_dtype == torch.uint8), int(self.clip_rewards),
                       self.w_obs.data_ptr(), self.w_actions.data_ptr(), self.w_rewards.data_ptr(),
                       self.w_to_play.data_ptr(), self.w_root_values.data_ptr(), self.w_child_visits.data_ptr())

  @_lib.on_device
  def open_chunk(self, capacity, obs_dtype=torch.float32):
    """Reserves `capacity` consecutive window positions for a history the self-play driver is still writing
    (`append_steps`); they stay reserved until `commit_chunk`.  Returns (chunk id, first position)."""
    if self.w_obs is None:
      self._obs_dtype = obs_dtype
      self.w_obs = torch.zeros((self.P, self.obs_elems), dtype=obs_dtype, device=self.device)
    elif obs_dtype != self._obs_dtype:
      raise TypeError("observations changed dtype between histories")
    start = self._alloc(int(capacity))
    cid = self._next_chunk
    self._next_chunk += 1
    self._register(cid, start, int(capacity))
    self._live[cid] = 1 << 30  # open: never recycled
    return cid, start

  @_lib.on_device
  def append_steps(self, dst_pos, obs, actions, rewards, to_play, root_values, child_visits):
    """One step of G games (device tensors, see mz_window_append): game g's record -> window position dst_pos[g]."""
    G = int(dst_pos.shape[0])
    if getattr(self, '_win_append', None) is None or self._win_append[0] != self.w_obs.data_ptr():
      self._win_append = (self.w_obs.data_ptr(), self._window_struct())
    _lib.check(self.lib.mz_window_append(self._win_append[1], G, _lib.ptr(dst_pos), _lib.ptr(obs), _lib.ptr(actions),
                                         _lib.ptr(rewards), _lib.ptr(to_play), _lib.ptr(root_values),
                                         _lib.ptr(child_visits), _lib.current_stream()), "mz_window_append")

  @_lib.on_device
  def copy_positions(self, src, dst, n):
    """Runs of window positions src[r].. -> dst[r].. (n[r] each): the overlap a running game's next chunk repeats."""
    if len(src) == 0:
      return
    dev = self.device
    t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a, dt)).to(dev)
    d_src, d_dst, d_n = t(src, np.int64), t(dst, np.int64), t(n, np.int32)
    if getattr(self, '_win_append', None) is None or self._win_append[0] != self.w_obs.data_ptr():
      self._win_append = (self.w_obs.data_ptr(), self._window_struct())
    _lib.check(self.lib.mz_window_copy(self._win_append[1], len(src), _lib.ptr(d_src), _lib.ptr(d_dst), _lib.ptr(d_n),
                                       _lib.current_stream()), "mz_window_copy")

  @_lib.on_device
  def commit_chunk(self, cid, start, n, errors, ignore=None, terminal=False):
    """save_history (replay_buffer.py:113-122) for a history that already sits in the window at [start, start + n):
    priorities from its `errors` (the last `ignore` steps of a running game get none), sum-tree slots, throughput."""
    errors = np.asarray(errors, np.float64)[:n]
    if ignore is not None:
      errors = errors[:-ignore] if ignore else errors
      priorities = self.get_priorities(errors) if len(errors) else []
    else:
      priorities = self.get_priorities(errors)
    # shrink the reservation to the history's real length
    self._chunk_at[cid][1] = n
    self._live[cid] = len(priorities)
    if n:
      for old, k in self.index.add(priorities, cid, start, n).items():
        if old in self._live:
          self._live[old] -= k
    self.throughput['frames'] += len(priorities)
    if terminal:
      self.throughput['games'] += 1

  @_lib.on_device
  def commit_chunks(self, chunks):
    """`commit_chunk` for every history the games of one self-play move finished, in order, with one sum-tree call:
    chunks = [(chunk id, start, n, errors, ignore, terminal)]."""
    items = []
    for cid, start, n, errors, ignore, terminal in chunks:
      errors = np.asarray(errors, np.float64)[:n]
      if ignore:
        errors = errors[:-ignore]
      priorities = self.get_priorities(errors) if len(errors) else np.zeros(0)
      self._chunk_at[cid][1] = n
      self._live[cid] = len(priorities)
      self.throughput['frames'] += len(priorities)
      if terminal:
        self.throughput['games'] += 1
      if n and len(priorities):
        items.append((priorities, cid, start, n))
    for old, k in self.index.add_chunks(items).items():
      if old in self._live:
        self._live[old] -= k

  @_lib.on_device
  def sample_batch(self):
    """replay_buffer.py:124-163 with numpy outputs like the reference: same draws from `random` and
    `np.random` in the same order (one random() per row, one randint per padded action)."""
    B, K, A = self.batch_size, self.num_unroll_steps, self.action_space
    self._step_beta()
    u01 = [random.random() for _ in range(B)]  # random.uniform(s1, s2) draws exactly one random()
    idx, pri, pos, cstart, clen, _ = self.index.sample(u01, self.beta, with_weights=False)
    h = self.index.last_blob.cpu().numpy()  # one device-to-host copy: idx | pos | chunk start | priority | - | len
    h_idx, h_pos, h_cs = h[0:B], h[B:2 * B], h[2 * B:3 * B]
    priorities, h_cl = h[3 * B:4 * B].view(np.float64), h[5 * B:6 * B].view(np.int32)[:B]
    # padding actions are drawn only where the slice is short (replay_buffer.py:149-152)
    n_real = np.clip(h_cl - (h_pos - h_cs), 0, K)
    pads = np.zeros((B, K), np.int32)
    for b in np.nonzero(n_real < K)[0]:
      for j in range(K - int(n_real[b])):
        pads[b, j] = np.random.randint(A)
    out = self._targets(pos, cstart, clen, torch.from_numpy(pads).to(self.device), False)
    hb = self._tgt_blob.cpu().numpy()  # one copy for the five outputs
    host, off = [], 0
    for i, (sh, n) in enumerate(self._tgt_layout):
      v = hb[off:off + int(np.prod(sh))]
      host.append((v.view(np.int32) if i == 1 else v).reshape(sh))
      off += n
    obs, actions, t_rewards, t_values, t_policies = host
    obs = obs.reshape((B,) + self.obs_space)
    sampling_probabilities = priorities / self.index.total_priority
    is_weights = np.power(self.index.num_memories * sampling_probabilities, -self.beta)
    is_weights /= is_weights.max()
    batch = (obs, actions.tolist(), (t_rewards, t_values, t_policies))
    return batch, h_idx.tolist(), is_weights

  @_lib.on_device
  def sample_batch_device(self, fuse_supports=True, ring=0):
    """Same sampling with nothing leaving the GPU and no host synchronisation: returns
    ((obs, actions [B,K] i32, t_rewards, t_values, t_policies[, value_support, reward_support]),
    idxs (int64 CUDA tensor, accepted by `update`), is_weights (float64 CUDA tensor)).  The padding
    actions are drawn for every row up front (one `np.random.randint(A, size=(B, K))`), which
    consumes the numpy stream differently from the reference; the sampled rows are the same.

    ring = R > 0: the outputs live in R preallocated sets used in turn (a batch stays valid until R - 1 further
    calls); the B draws of `random.random()` cross as the raw generator words they are made of (one
    `random.getrandbits`, same stream position afterwards, same float64 values formed on the device), the padding
    actions are drawn on the device from one np.random seed per batch, and sampling + target construction are ONE
    call into the library (`mz_replay_sample_targets`) with prebuilt arguments."""
    B, K, A = self.batch_size, self.num_unroll_steps, self.action_space
    self._step_beta()
    if ring:
      return self._sample_ring(fuse_supports, int(ring))
    u01 = [random.random() for _ in range(B)]
    idx, pri, pos, cstart, clen, isw = self.index.sample(u01, self.beta, with_weights=True)
    pads = torch.from_numpy(np.random.randint(A, size=(B, K)).astype(np.int32)).to(self.device)
    out = self._targets(pos, cstart, clen, pads, fuse_supports)
    return tuple(out), idx, isw

  def _sample_ring(self, fuse_supports, R):
    B, K, A = self.batch_size, self.num_unroll_steps, self.action_space
    key = (self.w_obs.data_ptr(), fuse_supports, R)
    st = getattr(self, '_ring', None)
    if st is None or st['key'] != key:
      st = self._ring = {'key': key, 'i': 0, 'sets': [self._ring_set(fuse_supports) for _ in range(R)]}
    e = st['sets'][st['i']]
    st['i'] = (st['i'] + 1) % R
    e['event'].synchronize()  # the copy that last read this set's pinned blob is done (it long is)
    # B draws of random.random() (random.uniform(s1, s2) consumes exactly one each) as the 2 * B raw generator
    # outputs they are made of; the sampling kernel forms the float64 values (mz_replay_sample_targets)
    C.memmove(e['h_ptr'], random.getrandbits(64 * B).to_bytes(8 * B, 'little'), 8 * B)
    e['d_in'].copy_(e['h_in'], non_blocking=True)
    e['event'].record()
    args = e['args']
    args[7], args[8], args[26] = self.index.ring.num_memories, float(self.beta), torch.cuda.current_stream().cuda_stream
    if not st.get('seeds'):  # seeds of the device-drawn padding actions: 256 np.random draws at a time
      st['seeds'] = np.random.randint(1, 1 << 62, size=256).tolist()
    args[18] = st['seeds'].pop()
    rc = self.lib.mz_replay_sample_targets(*args)
    if rc:
      _lib.check(rc, "mz_replay_sample_targets")
    return e['ret']

  def _ring_set(self, fuse_supports):
    """One preallocated set of sampling + target outputs, its pinned input blob and the argument list of
    mz_replay_sample_targets."""
    B, K, A = self.batch_size, self.num_unroll_steps, self.action_space
    dev, idx = self.device, self.index
    h_in = torch.zeros(8 * B, dtype=torch.uint8).pin_memory()
    d_in = torch.zeros_like(h_in, device=dev)
    d_u = d_in.view(torch.float64)
    d_pads = torch.zeros(B * max(K, 1), dtype=torch.int32, device=dev)
    blob = torch.zeros(6 * B, dtype=torch.int64, device=dev)
    t_idx, pos, cstart = blob[0:B], blob[B:2 * B], blob[2 * B:3 * B]
    pri, isw = blob[3 * B:4 * B].view(torch.float64), blob[4 * B:5 * B].view(torch.float64)
    clen = blob[5 * B:6 * B].view(torch.int32)[:B]
    vb = self.value_support[1] - self.value_support[0] + 1
    rb = self.reward_support[1] - self.reward_support[0] + 1
    shapes = [(B, self.obs_elems), (B, K), (B, K + 1), (B, K + 1), (B, K + 1, A)]
    if fuse_supports:
      shapes += [(B, K + 1, vb), (B, K + 1, rb)]
    sizes = [(int(np.prod(sh)) + 3) // 4 * 4 for sh in shapes]
    tb = torch.zeros(sum(sizes), dtype=torch.float32, device=dev)
    out, off = [], 0
    for i, (sh, n) in enumerate(zip(shapes, sizes)):
      v = tb[off:off + int(np.prod(sh))]
      out.append((v.view(torch.int32) if i == 1 else v).view(sh))
      off += n
    win = self._window_struct()
    cfg = _lib.TargetCfg(B, K, self.td_steps, int(fuse_supports), self.value_support[0], self.value_support[1],
                         self.reward_support[0], self.reward_support[1], int(self.no_target_transform), 0,
                         float(self.discount**self.td_steps), self.d_discounts.data_ptr(), None, None)
    P = lambda t: C.c_void_p(t.data_ptr())
    outs = [P(t) for t in out] + ([None, None] if not fuse_supports else [])
    args = [P(idx.tree), idx.max_capacity, P(d_u), 1, P(idx.slot_pos), P(idx.slot_start), P(idx.slot_len), 0, 0.0,
            P(t_idx), P(pri), P(pos), P(cstart), P(clen), P(isw), win, cfg, P(d_pads), 1] + outs + [None]
    return {'h_in': h_in, 'd_in': d_in, 'h_ptr': h_in.data_ptr(), 'event': torch.cuda.Event(), 'args': args,
            'ret': (tuple(out), t_idx, isw), 'keep': (blob, tb, win, cfg, d_pads)}

  @_lib.on_device
  def update(self, idxs, errors):
    """replay_buffer.py:200-203.  `idxs` may be the list `sample_batch` returned or the CUDA tensor
    of `sample_batch_device`; `errors` a numpy array (learners.py:183) or a CUDA tensor.  With both on the device,
    float32 errors and alpha == 1 (the reference's default) nothing comes back to the host: the priorities are
    computed by the library (`mz_sumtree_update_errors`, bit-exact for alpha == 1)."""
    if (torch.is_tensor(errors) and torch.is_tensor(idxs) and errors.is_cuda and idxs.is_cuda and
        errors.dtype == torch.float32 and idxs.dtype == torch.int64 and float(self.alpha) == 1.0):
      n = int(idxs.shape[0])
      if n == 0:
        return
      errors, idxs = errors.detach().contiguous(), idxs.contiguous()
      ws = getattr(self, '_upd_ws', None)
      if ws is None or ws.shape[1] < n:
        ws = self._upd_ws = torch.empty((2, n), dtype=torch.float64, device=self.device)
      rc = self.lib.mz_sumtree_update_errors(C.c_void_p(self.index.tree.data_ptr()), self.index.max_capacity, n,
                                             C.c_void_p(idxs.data_ptr()), C.c_void_p(errors.data_ptr()),
                                             float(self.epsilon), 1.0, C.c_void_p(ws[0].data_ptr()),
                                             C.c_void_p(ws[1].data_ptr()), torch.cuda.current_stream().cuda_stream)
      if rc:
        _lib.check(rc, "mz_sumtree_update_errors")
      return
    if torch.is_tensor(errors):
      errors = errors.detach().cpu().numpy()
    if torch.is_tensor(idxs):
      idxs = idxs.cpu().numpy()
    priorities = self.get_priorities(np.asarray(errors))
    self.index.update(np.asarray(idxs, np.int64), priorities)

  def size(self):
    return self.index.num_memories

  def get_throughput(self):
    return self.throughput

  # -- internals -----------------------------------------------------------------------------------
  def _step_beta(self):
    if self.index.num_memories == 0:
      raise _lib.MzError("sample_batch on an empty replay buffer")
    if self.beta < 1:  # replay_buffer.py:137: np.min([1., beta + increment]) -- same float64 arithmetic, without the array
      self.beta = min(1., self.beta + self.beta_increment_per_sampling)

  def _targets(self, d_pos, d_cs, d_cl, d_pads, fuse_supports):
    B, K, A = self.batch_size, self.num_unroll_steps, self.action_space
    dev = self.device
    vb = self.value_support[1] - self.value_support[0] + 1
    rb = self.reward_support[1] - self.reward_support[0] + 1
    # one allocation per batch, carved into the output tensors (all float32 / int32: 4-byte elements)
    shapes = [(B, self.obs_elems), (B, K), (B, K + 1), (B, K + 1), (B, K + 1, A)]
    if fuse_supports:
      shapes += [(B, K + 1, vb), (B, K + 1, rb)]
    sizes = [(int(np.prod(sh)) + 3) // 4 * 4 for sh in shapes]  # every view starts 16-byte aligned
    blob = torch.empty(sum(sizes), dtype=torch.float32, device=dev)
    out, off = [], 0
    for i, (sh, n) in enumerate(zip(shapes, sizes)):
      v = blob[off:off + int(np.prod(sh))]
      out.append((v.view(torch.int32) if i == 1 else v).view(sh))
      off += n
    if getattr(self, '_tgt_structs', None) is None or self._tgt_structs[0] != (self.w_obs.data_ptr(), fuse_supports):
      win = _lib.Window(A, self.obs_elems, int(self._obs_dtype == torch.uint8), int(self.clip_rewards),
                        self.w_obs.data_ptr(), self.w_actions.data_ptr(), self.w_rewards.data_ptr(),
                        self.w_to_play.data_ptr(), self.w_root_values.data_ptr(),
                        self.w_child_visits.data_ptr())
      cfg = _lib.TargetCfg(B, K, self.td_steps, int(fuse_supports), self.value_support[0],
                           self.value_support[1], self.reward_support[0], self.reward_support[1],
                           int(self.no_target_transform), 0, float(self.discount**self.td_steps),
                           self.d_discounts.data_ptr(), None, None)
      self._tgt_structs = ((self.w_obs.data_ptr(), fuse_supports), win, cfg)
    _, win, cfg = self._tgt_structs
    self._tgt_blob, self._tgt_layout = blob, list(zip(shapes, sizes))
    ptrs = [_lib.ptr(t) for t in out] + ([None, None] if not fuse_supports else [])
    _lib.check(self.lib.mz_build_targets(win, cfg, _lib.ptr(d_pos), _lib.ptr(d_cs), _lib.ptr(d_cl),
                                         _lib.ptr(d_pads), *ptrs, _lib.current_stream()),
               "mz_build_targets")
    return out

  def _alloc(self, n):
    """Ring allocation of n consecutive window positions.  A chunk stays readable as long as a
    sum-tree slot refers to it (the reference keeps the history object alive through
    SumTree.buffer); chunks without slots are recycled when the ring comes round."""
    if n > self.P:
      raise _lib.MzError("history of %d steps does not fit the replay window arena (%d)" % (n, self.P))
    wraps = 0
    while True:
      if self._head + n > self.P:
        self._head = 0
        wraps += 1
        if wraps > 2:
          raise _lib.MzError("replay window arena is full of live histories; construct "
                             "PrioritizedReplay with a larger window_positions")
      lo, hi = self._head, self._head + n
      hit = self._overlapping(lo, hi)
      blocker = [self._chunk_at[c][0] + self._chunk_at[c][1] for c in hit if self._live.get(c, 0) > 0]
      if not blocker:
        break
      self._head = max(blocker)  # a history that is still sampled (or still being written): allocate behind it
    for cid in hit:
      self._live.pop(cid, None)
      start = self._chunk_at.pop(cid)[0]
      del self._starts[bisect.bisect_left(self._starts, (start, cid))]
    self._head = hi
    return lo

  def _register(self, cid, start, n):
    self.max_chunk = max(self.max_chunk, n)
    self._chunk_at[cid] = [start, n]
    bisect.insort(self._starts, (start, cid))

  def _overlapping(self, lo, hi):
    """Chunk ids whose reservation [start, start + length) meets [lo, hi) (chunks never overlap one another)."""
    out = []
    i = bisect.bisect_left(self._starts, (hi, -1)) - 1
    while i >= 0:
      start, cid = self._starts[i]
      ln = self._chunk_at[cid][1]
      if start + self.max_chunk <= lo:  # no earlier reservation is long enough to reach `lo`
        break
      if start < hi and start + ln > lo:
        out.append(cid)
      i -= 1
    return out

  def _upload(self, history, n):
    obs = np.stack([np.asarray(o) for o in history.observations[:n]]).reshape(n, -1)
    dt = torch.uint8 if obs.dtype == np.uint8 else torch.float32
    if self.w_obs is None:
      self._obs_dtype = dt
      self.w_obs = torch.zeros((self.P, self.obs_elems), dtype=dt, device=self.device)
    elif dt != self._obs_dtype:
      raise TypeError("observations changed dtype between histories")
    if obs.shape[1] != self.obs_elems:
      raise ValueError("observation has %d elements, config.obs_space says %d" % (obs.shape[1], self.obs_elems))
    start = self._alloc(n)
    sl = slice(start, start + n)
    dev = self.device
    self.w_obs[sl] = torch.from_numpy(np.ascontiguousarray(obs if dt == torch.uint8 else obs.astype(np.float32))).to(dev)
    self.w_actions[sl] = torch.from_numpy(np.asarray(history.actions, np.int32)).to(dev)
    self.w_rewards[sl] = torch.from_numpy(np.asarray(history.rewards, np.float64).astype(np.float32)).to(dev)
    self.w_to_play[sl] = torch.from_numpy(np.asarray(history.to_play, np.int8)).to(dev)
    self.w_root_values[sl] = torch.from_numpy(np.asarray(history.root_values, np.float64)).to(dev)
    self.w_child_visits[sl] = torch.from_numpy(
        np.asarray(history.child_visits, np.float64).reshape(n, self.action_space).astype(np.float32)).to(dev)
    return start
