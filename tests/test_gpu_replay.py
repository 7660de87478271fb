"""GPU parity tests of the PrioritizedReplay facade (device window + device sum-tree + fused target
kernel) against goldens produced by the unmodified reference (replay_buffer.py), and of the
sum-tree kernels against the CPU oracle at larger sizes."""
import collections
import random
import types

import numpy as np
import pytest
import torch

from helpers import REPLAY_CASES, is_clip_case, load

pytestmark = pytest.mark.gpu

HistorySlice = collections.namedtuple(
    "HistorySlice", "observations child_visits root_values actions rewards errors dones steps "
    "env_states to_play")


def _config(g, **kw):
  d = dict(batch_size=int(g["batch_size"]), beta_increment_per_sampling=float(g["beta_increment"]),
           epsilon=float(g["epsilon"]), alpha=float(g["alpha"]), beta=float(g["beta"]),
           stored_before_train=0, num_unroll_steps=int(g["num_unroll_steps"]),
           td_steps=int(g["td_steps"]), discount=float(g["discount"]),
           action_space=int(g["action_space"]), obs_space=(int(g["obs_dim"]),),
           window_size=int(g["window_size"]), window_step=int(g["window_step"]), seed=None,
           value_support=(-15, 15), reward_support=(-15, 15), no_target_transform=False,
           clip_rewards=is_clip_case(g))
  d.update(kw)
  return types.SimpleNamespace(**d)


def _history(g, h):
  n = len(g["h%d_root_values" % h])
  rewards = g["h%d_rewards" % h].tolist()
  if is_clip_case(g):  # raw environment rewards; integral ones as python ints, like gym returns them
    rewards = [int(r) if r == int(r) and abs(r) >= 1 else r for r in rewards]
  return HistorySlice([o for o in g["h%d_obs" % h]], g["h%d_child_visits" % h].tolist(),
                      g["h%d_root_values" % h].tolist(), g["h%d_actions" % h].tolist(),
                      rewards, g["h%d_errors" % h].tolist(), [False] * n,
                      list(range(n)), [None] * n, g["h%d_to_play" % h].tolist())


def _filled(g):
  from model_based_rl_b200.replay_buffer import PrioritizedReplay
  rb = PrioritizedReplay(_config(g))
  for h in range(int(g["n_hist"])):
    ign = int(g["ignores"][h])
    rb.save_history(_history(g, h), ignore=None if ign < 0 else ign, terminal=ign < 0)
  return rb


class _Feed(object):
  """Replaces the two random sources of sample_batch with the recorded draws of the golden run."""

  def __init__(self, frac, pads, n_real):
    self.frac, self.i = list(frac), 0
    # the reference draws pads row by row, only for the missing tail of a short action slice
    self.pad_stream = [int(pads[b, k]) for b in range(len(pads)) for k in range(pads.shape[1] - n_real[b])]
    self.j = 0

  def random(self):
    self.i += 1
    return self.frac[self.i - 1]

  def randint(self, n):
    self.j += 1
    return self.pad_stream[self.j - 1]


@pytest.mark.parametrize("case", REPLAY_CASES)
def test_facade_matches_reference(case, monkeypatch):
  from model_based_rl_b200 import replay_buffer as rbmod
  g = load("replay_" + case)
  B, K = int(g["batch_size"]), int(g["num_unroll_steps"])
  rb = _filled(g)
  lens = [len(g["h%d_root_values" % h]) for h in range(int(g["n_hist"]))]
  assert rb.size() == int(g["b0_num_memories"])
  for b in range(int(g["n_batches"])):
    assert rb.index.total_priority == float(g["b%d_total_priority" % b])
    n_real = np.clip(np.array([lens[h] for h in g["b%d_hist" % b]]) - g["b%d_steps" % b], 0, K)
    feed = _Feed(g["b%d_frac" % b], g["b%d_pads" % b], n_real)
    monkeypatch.setattr(rbmod.random, "random", feed.random)
    monkeypatch.setattr(np.random, "randint", feed.randint)
    (obs, actions, (t_r, t_v, t_p)), idxs, is_w = rb.sample_batch()
    monkeypatch.undo()
    assert feed.i == B and feed.j == len(feed.pad_stream)
    assert idxs == g["b%d_idxs" % b].tolist()                       # same rows: bit-exact tree walk
    assert rb.beta == float(g["b%d_beta_after" % b])
    assert np.array_equal(is_w, g["b%d_is_weights" % b]) and is_w.dtype == np.float64
    assert np.array_equal(obs, g["b%d_obs" % b]) and obs.dtype == np.float32
    assert actions == g["b%d_actions" % b].tolist()
    assert np.array_equal(t_r, g["b%d_t_rewards" % b])
    assert np.array_equal(t_p, g["b%d_t_policies" % b])
    want = g["b%d_t_values" % b]
    assert np.max(np.abs(t_v - want) / np.maximum(np.abs(want), 1.0)) <= 1e-5  # north_star tolerance
    rb.update(idxs, g["b%d_update_errors" % b])
    torch.cuda.synchronize()
    assert np.array_equal(rb.index.tree.cpu().numpy(), g["b%d_tree_after_update" % b])  # bit-exact sums


@pytest.mark.parametrize("ring", [0, 3], ids=["fresh_outputs", "ring_of_3"])
@pytest.mark.parametrize("case", REPLAY_CASES)
def test_device_sampling_matches_reference(case, ring, monkeypatch):
  """sample_batch_device: same rows, importance weights from the device pow (<= 1e-14 relative),
  supports fused; idxs/errors may stay CUDA tensors for update() (float32 errors and alpha == 1: priorities
  computed on the device, tree bit-exact against the reference's).  ring = 3: preallocated output sets, pinned
  staging, one library call per batch (mz_replay_sample_targets)."""
  from model_based_rl_b200 import replay_buffer as rbmod
  from model_based_rl_b200.config import Config
  g = load("replay_" + case)
  B, K = int(g["batch_size"]), int(g["num_unroll_steps"])
  rb = _filled(g)
  cfg = Config(dict(value_support=[-15, 15], reward_support=[-15, 15], no_target_transform=False))
  for b in range(int(g["n_batches"])):
    frac = list(g["b%d_frac" % b])
    pads = g["b%d_pads" % b]
    monkeypatch.setattr(rbmod.random, "random", lambda frac=frac: frac.pop(0))
    if ring:
      # the ring path takes the B draws as raw MT19937 words (random.getrandbits) and forms the float64 values on
      # the device: feed it the words the golden fractions are made of, random() = ((a >> 5) * 2**26 + (b >> 6)) / 2**53
      def words(k, frac=list(frac)):
        assert k == 64 * B
        x = 0
        for i, f in enumerate(frac):
          n = int(f * 2**53)
          assert n / 2**53 == f
          x |= (((n >> 26) << 5) | ((n & (2**26 - 1)) << 6 << 32)) << (64 * i)
        return x
      monkeypatch.setattr(rbmod.random, "getrandbits", words)
    else:
      monkeypatch.setattr(np.random, "randint", lambda A, size=None, pads=pads, **kw: pads)
    out, idxs, is_w = rb.sample_batch_device(fuse_supports=True, ring=ring)
    monkeypatch.undo()
    assert idxs.is_cuda and idxs.cpu().tolist() == g["b%d_idxs" % b].tolist()
    want_w = g["b%d_is_weights" % b]
    assert is_w.dtype == torch.float64
    assert np.max(np.abs(is_w.cpu().numpy() - want_w) / want_w) <= 1e-14
    obs, actions, t_r, t_v, t_p, vs, rs = out
    assert np.array_equal(obs.cpu().numpy(), g["b%d_obs" % b])
    if ring:  # padding actions drawn on the device: the real prefix of every action slice is the reference's
      lens = [len(g["h%d_root_values" % h]) for h in range(int(g["n_hist"]))]
      n_real = np.clip(np.array([lens[h] for h in g["b%d_hist" % b]]) - g["b%d_steps" % b], 0, K)
      got, want_a = actions.cpu().numpy(), g["b%d_actions" % b]
      real = np.arange(K)[None, :] < n_real[:, None]
      assert np.array_equal(got[real], want_a[real])
      assert ((got >= 0) & (got < int(g["action_space"]))).all()
    else:
      assert np.array_equal(actions.cpu().numpy(), g["b%d_actions" % b])
    assert np.array_equal(t_r.cpu().numpy(), g["b%d_t_rewards" % b])
    assert torch.equal(vs, cfg.value_phi(Config.scalar_transform(t_v)))
    assert torch.equal(rs, cfg.reward_phi(Config.scalar_transform(t_r)))
    rb.update(idxs, torch.from_numpy(g["b%d_update_errors" % b]).cuda())
    torch.cuda.synchronize()
    assert np.array_equal(rb.index.tree.cpu().numpy(), g["b%d_tree_after_update" % b])


def test_device_priority_update_equals_host_update():
  """update() with CUDA idxs + float32 CUDA errors (mz_sumtree_update_errors) against update() with the same values
  as numpy arrays (numpy's (|e| + eps) ** alpha on the host): identical float64 trees, repeated indices included."""
  from model_based_rl_b200.replay_buffer import PrioritizedReplay
  rng = np.random.default_rng(3)
  cfg = types.SimpleNamespace(batch_size=64, beta_increment_per_sampling=0.001, epsilon=0.01, alpha=1.0, beta=0.5,
                              num_unroll_steps=3, td_steps=5, discount=0.997, action_space=4, obs_space=(8,),
                              window_size=3000, window_step=None, seed=None)
  a, b = PrioritizedReplay(cfg), PrioritizedReplay(cfg)
  for rb in (a, b):
    r2 = np.random.default_rng(5)
    for _ in range(12):
      n = int(r2.integers(50, 400))
      h = HistorySlice(list(r2.normal(size=(n, 8)).astype(np.float32)), r2.random((n, 4)).tolist(), r2.normal(size=n).tolist(),
                       r2.integers(0, 4, size=n).tolist(), r2.normal(size=n).tolist(), np.abs(r2.normal(size=n)).tolist(),
                       [False] * n, list(range(n)), [None] * n, [1] * n)
      rb.save_history(h, ignore=None, terminal=True)
  leaf0 = a.index.max_capacity - 1
  for _ in range(10):
    idx = leaf0 + rng.integers(0, a.size(), size=512)
    idx[100:110] = idx[0]  # the same memory sampled several times in one batch: the last update wins
    err = rng.normal(size=512).astype(np.float32)
    a.update(idx.tolist(), err)
    b.update(torch.from_numpy(idx).cuda(), torch.from_numpy(err).cuda())
  torch.cuda.synchronize()
  assert np.array_equal(a.index.tree.cpu().numpy(), b.index.tree.cpu().numpy())
  assert a.index.total_priority == b.index.total_priority > 0


@pytest.mark.parametrize("capacity,step", [(1000, 250), (4096, 4096), (200_000, 200_000)])
def test_sumtree_kernels_match_oracle(capacity, step):
  """Random adds / batched updates with repeated leaves, capacities that are not powers of two
  (leaves on two depths), the Breakout window size: device sums == oracle sums bit for bit."""
  from oracle import replay_ref
  from model_based_rl_b200.replay_buffer import ReplayIndex
  rng = np.random.default_rng(capacity)
  dev = torch.device("cuda", 0)
  idx = ReplayIndex(capacity, step, dev)
  ref = replay_ref.SumTreeRef(capacity, step)
  start = 0
  for chunk in range(12):
    n = int(rng.integers(1, min(capacity, 5000)))
    pri = rng.random(n) ** 3 + 1e-3
    idx.add(pri, chunk, start, n + 7)
    ref.add(pri, chunk)
    start += n + 7
  torch.cuda.synchronize()
  assert np.array_equal(idx.tree.cpu().numpy(), ref.tree)
  assert idx.num_memories == ref.num_memories
  B = 512
  for it in range(4):
    u = rng.random(B)
    d_idx, d_pri, d_pos, d_cs, d_cl, d_w = idx.sample(u, 0.4 + 0.1 * it, with_weights=True)
    picks, w = replay_ref.sample_indices(ref, B, u, 0.4 + 0.1 * it)
    assert d_idx.cpu().tolist() == [p[0] for p in picks]
    assert np.array_equal(d_pri.cpu().numpy(), np.array([p[1] for p in picks]))
    assert np.array_equal((d_pos - d_cs).cpu().numpy(), np.array([p[2] for p in picks]))
    assert np.max(np.abs(d_w.cpu().numpy() - w) / w) <= 1e-14
    # batched update with repeated leaves (stratified sampling can hit one leaf several times)
    upd = np.concatenate([np.array([p[0] for p in picks]), np.array([p[0] for p in picks[:40]])])
    pri = rng.random(len(upd)) + 1e-3
    idx.update(upd, pri)
    for i, p in zip(upd, pri):
      ref.update(int(i), p)
    torch.cuda.synchronize()
    assert np.array_equal(idx.tree.cpu().numpy(), ref.tree)


def test_window_arena_recycles_dead_chunks():
  """A small ring: old chunks are overwritten slot by slot, their window positions get recycled,
  and every sampled row still reproduces the oracle's targets for the chunk it points at."""
  import oracle
  from model_based_rl_b200.replay_buffer import PrioritizedReplay
  rng = np.random.default_rng(11)
  A, K, T, E = 4, 5, 10, 8
  cfg = types.SimpleNamespace(
      batch_size=64, beta_increment_per_sampling=0.001, epsilon=0.01, alpha=1.0, beta=0.5,
      num_unroll_steps=K, td_steps=T, discount=0.997, action_space=A, obs_space=(E,), window_size=300,
      window_step=None, seed=5, max_history_length=60)
  rb = PrioritizedReplay(cfg, window_positions=700)
  chunks = {}
  for c in range(40):
    n = int(rng.integers(20, 61))
    cv = rng.random((n, A))
    cv /= cv.sum(1, keepdims=True)
    h = HistorySlice([rng.normal(size=E).astype(np.float32) for _ in range(n)], cv.tolist(),
                     rng.normal(0, 2, size=n).tolist(), rng.integers(0, A, size=n).tolist(),
                     np.sign(rng.normal(size=n)).tolist(), rng.normal(size=n).tolist(), [False] * n,
                     list(range(n)), [None] * n, [1] * n)
    rb.save_history(h, terminal=True)
    chunks[c] = h
    assert rb.size() == min(300, sum(len(x.root_values) for x in chunks.values()))
    (obs, actions, (t_r, t_v, t_p)), idxs, is_w = rb.sample_batch()
    slots = np.array(idxs) - (300 - 1)
    for b in range(64):
      cid = int(rb.index.slot_chunk[slots[b]])
      hh = chunks[cid]
      step = int(np.argmax([np.array_equal(o, obs[b]) for o in hh.observations]))
      assert np.array_equal(hh.observations[step], obs[b])
      r, v, p = oracle.insert_target(np.array(hh.rewards), np.array(hh.to_play, np.int8),
                                     np.array(hh.root_values), np.array(hh.child_visits), K, T, 0.997,
                                     step)
      assert np.array_equal(t_r[b], r) and np.array_equal(t_p[b], p)
      assert np.allclose(t_v[b], v, rtol=1e-5, atol=1e-6)


def test_history_longer_than_the_ring_laps_it_like_the_reference():
  """A growing ring (window_step) shorter than one history: a single save_history wraps it, the same slot
  is written more than once inside one add.  The reference writes one memory at a time (replay_buffer.py:
  19-33): the last write to a slot wins, every earlier one is an ordinary overwrite.  Sums, slot -> (history,
  step) mapping, num_memories and the liveness of the window chunks must all follow that."""
  from oracle import replay_ref
  from model_based_rl_b200.replay_buffer import PrioritizedReplay
  rng = np.random.default_rng(3)
  A, K, T, E = 3, 2, 3, 4
  cfg = types.SimpleNamespace(
      batch_size=32, beta_increment_per_sampling=0.0, epsilon=0.01, alpha=0.7, beta=0.5,
      num_unroll_steps=K, td_steps=T, discount=0.99, action_space=A, obs_space=(E,), window_size=40,
      window_step=8, seed=7, max_history_length=64)
  rb = PrioritizedReplay(cfg, window_positions=400)
  ref = replay_ref.SumTreeRef(40, 8)
  hist = {}
  for c, n in enumerate([5, 30, 9, 70, 3, 41, 12, 100, 7]):
    cv = rng.random((n, A))
    cv /= cv.sum(1, keepdims=True)
    h = HistorySlice([rng.normal(size=E).astype(np.float32) for _ in range(n)], cv.tolist(),
                     rng.normal(0, 2, size=n).tolist(), rng.integers(0, A, size=n).tolist(),
                     rng.normal(size=n).tolist(), rng.normal(size=n).tolist(), [False] * n,
                     list(range(n)), [None] * n, [1] * n)
    rb.save_history(h, terminal=True)
    ref.add(replay_ref.get_priorities(np.array(h.errors), 0.01, 0.7), c)
    hist[c] = h
    torch.cuda.synchronize()
    assert np.array_equal(rb.index.tree.cpu().numpy(), ref.tree)
    assert rb.size() == ref.num_memories
    live_slots = ref.slot_hist >= 0
    assert np.array_equal(rb.index.slot_chunk[live_slots], ref.slot_hist[live_slots])
    pos = rb.index.slot_pos.cpu().numpy()
    start = rb.index.slot_start.cpu().numpy()
    assert np.array_equal((pos - start)[live_slots], ref.slot_step[live_slots])
    # liveness: a chunk is referenced by exactly the slots that still point at it
    for cid, n_live in rb._live.items():
      assert n_live == int((ref.slot_hist == cid).sum()), (cid, n_live)
    # every sampled row reads the observation of the (history, step) the reference's slot holds
    (obs, actions, targets), idxs, is_w = rb.sample_batch()
    for b, ti in enumerate(idxs):
      slot = ti - (40 - 1)
      hh = hist[int(ref.slot_hist[slot])]
      assert np.array_equal(obs[b], hh.observations[int(ref.slot_step[slot])])


@pytest.mark.parametrize("capacity,step", [(700, 200), (5000, 5000)])
def test_add_chunks_equals_one_add_per_chunk(capacity, step):
  """ReplayIndex.add_chunks (every chunk a self-play move finished: one upload, one mz_sumtree_add_chunks call)
  against one ReplayIndex.add per chunk: identical float64 tree, identical slot -> (position, chunk start, chunk
  length) tables, identical ring cursor and overwritten-slot counts -- also across ring wraps and capacity growth."""
  from model_based_rl_b200.replay_buffer import ReplayIndex
  rng = np.random.default_rng(capacity)
  a, b = ReplayIndex(capacity, step, "cuda"), ReplayIndex(capacity, step, "cuda")
  cid, start = 0, 0
  for move in range(40):
    items = []
    for _ in range(int(rng.integers(1, 7))):
      n = int(rng.integers(1, 90))
      k = int(rng.integers(0, n + 1))  # a running game's chunk has fewer priorities than steps
      items.append((rng.random(k) + 0.01, cid, start, n))
      cid, start = cid + 1, start + n
    want = {}
    for pri, c, s, n in items:
      for old, cnt in a.add(pri, c, s, n).items():
        want[old] = want.get(old, 0) + cnt
    got = b.add_chunks(items)
    assert got == want, move
  torch.cuda.synchronize()
  assert a.ring.__dict__ == b.ring.__dict__
  assert np.array_equal(a.slot_chunk, b.slot_chunk)
  for name in ("tree", "slot_pos", "slot_start", "slot_len"):
    assert np.array_equal(getattr(a, name).cpu().numpy(), getattr(b, name).cpu().numpy()), name


def test_ring_sampling_consumes_random_like_the_reference_loop():
  """The ring path's one random.getrandbits(64 * B) leaves the `random` module in the state B calls of
  random.random() leave it in, and draws the same rows: a run that mixes both paths stays on the reference's stream."""
  from model_based_rl_b200.replay_buffer import PrioritizedReplay
  cfg = types.SimpleNamespace(batch_size=96, beta_increment_per_sampling=0.001, epsilon=0.01, alpha=1.0, beta=0.5,
                              num_unroll_steps=3, td_steps=5, discount=0.997, action_space=4, obs_space=(8,),
                              window_size=2000, window_step=None, seed=None)
  rb = PrioritizedReplay(cfg)
  r2 = np.random.default_rng(5)
  for _ in range(8):
    n = int(r2.integers(50, 300))
    rb.save_history(HistorySlice(list(r2.normal(size=(n, 8)).astype(np.float32)), r2.random((n, 4)).tolist(),
                                 r2.normal(size=n).tolist(), r2.integers(0, 4, size=n).tolist(), r2.normal(size=n).tolist(),
                                 np.abs(r2.normal(size=n)).tolist(), [False] * n, list(range(n)), [None] * n, [1] * n),
                    ignore=None, terminal=True)
  random.seed(123)
  _, idx_a, w_a = rb.sample_batch_device(True)
  state_a = random.getstate()
  idx_a, w_a = idx_a.cpu().numpy().copy(), w_a.cpu().numpy().copy()
  rb.beta = 0.5
  random.seed(123)
  _, idx_b, w_b = rb.sample_batch_device(True, ring=2)
  assert random.getstate() == state_a
  assert np.array_equal(idx_a, idx_b.cpu().numpy()) and np.array_equal(w_a, w_b.cpu().numpy())
