"""GPU parity tests for the target kernel and the scalar transform kernels against goldens produced
by the unmodified reference (replay_buffer.PrioritizedReplay.sample_batch, config.Config)."""
import ctypes as C
import types

import numpy as np
import pytest
import torch

import oracle
from helpers import REPLAY_CASES, is_clip_case, load

pytestmark = pytest.mark.gpu


def _window_from_golden(g):
  """Concatenates the golden chunks into the SoA replay window the kernel reads."""
  n_hist = int(g["n_hist"])
  starts, lens = [], []
  cur = 0
  cols = dict(obs=[], actions=[], rewards=[], to_play=[], root_values=[], child_visits=[])
  for h in range(n_hist):
    n = len(g["h%d_root_values" % h])
    starts.append(cur)
    lens.append(n)
    cols["obs"].append(g["h%d_obs" % h][:n])
    cols["actions"].append(g["h%d_actions" % h])
    cols["rewards"].append(g["h%d_rewards" % h].astype(np.float32))
    cols["to_play"].append(g["h%d_to_play" % h])
    cols["root_values"].append(g["h%d_root_values" % h])
    cols["child_visits"].append(g["h%d_child_visits" % h].astype(np.float32))
    cur += n
  dev = {k: torch.from_numpy(np.concatenate(v)).cuda().contiguous() for k, v in cols.items()}
  return dev, np.array(starts, np.int64), np.array(lens, np.int32)


@pytest.fixture(params=[0, 1, 2], ids=["kernel_by_shape", "warp_per_row_kernel", "lane_per_position_kernel"])
def targets_kernel(request):
  """The three kernels behind mz_build_targets (include/mzb200.h: TMA-staged by shape, warp per row, lane per
  position) run every case they are able to."""
  from model_based_rl_b200 import _lib
  lib = _lib.load()
  lib.mz_debug_set_targets_kernel(request.param)
  yield request.param
  lib.mz_debug_set_targets_kernel(0)


@pytest.mark.parametrize("case", REPLAY_CASES)
@pytest.mark.parametrize("fuse", [False, True])
def test_build_targets_matches_reference(case, fuse, targets_kernel):
  from model_based_rl_b200 import _lib
  from model_based_rl_b200.config import Config
  lib = _lib.load()
  g = load("replay_" + case)
  A, K, T, B = int(g["action_space"]), int(g["num_unroll_steps"]), int(g["td_steps"]), int(g["batch_size"])
  E, disc = int(g["obs_dim"]), float(g["discount"])
  dev, starts, lens = _window_from_golden(g)
  # the clip fixture holds raw rewards: the kernel clips on the fly (mz_window.clip_rewards, wrappers.py:236-238)
  win = _lib.Window(A, E, int(g["obs_uint8"]), int(is_clip_case(g)), dev["obs"].data_ptr(), dev["actions"].data_ptr(),
                    dev["rewards"].data_ptr(), dev["to_play"].data_ptr(),
                    dev["root_values"].data_ptr(), dev["child_visits"].data_ptr())
  discounts = torch.from_numpy(np.array([disc**n for n in range(K + T)], np.float32)).cuda()
  tc = _lib.TargetCfg(B, K, T, int(fuse), -15, 15, -15, 15, 0, 0, disc**T, discounts.data_ptr(), None, None)
  cfgobj = Config(dict(value_support=[-15, 15], reward_support=[-15, 15], no_target_transform=False))
  worst = 0.0
  for b in range(int(g["n_batches"])):
    hist, steps = g["b%d_hist" % b], g["b%d_steps" % b]
    pos = torch.from_numpy(starts[hist] + steps).cuda()
    cs = torch.from_numpy(starts[hist]).cuda()
    cl = torch.from_numpy(lens[hist]).cuda()
    pads = torch.from_numpy(g["b%d_pads" % b].astype(np.int32)).cuda()
    obs = torch.zeros((B, E), dtype=torch.float32, device="cuda")
    acts = torch.zeros((B, K), dtype=torch.int32, device="cuda")
    tr = torch.zeros((B, K + 1), dtype=torch.float32, device="cuda")
    tv = torch.zeros((B, K + 1), dtype=torch.float32, device="cuda")
    tp = torch.zeros((B, K + 1, A), dtype=torch.float32, device="cuda")
    vs = torch.zeros((B, K + 1, 31), dtype=torch.float32, device="cuda")
    rs = torch.zeros((B, K + 1, 31), dtype=torch.float32, device="cuda")
    _lib.check(lib.mz_build_targets(win, tc, _lib.ptr(pos), _lib.ptr(cs), _lib.ptr(cl), _lib.ptr(pads),
                                    _lib.ptr(obs), _lib.ptr(acts), _lib.ptr(tr), _lib.ptr(tv),
                                    _lib.ptr(tp), _lib.ptr(vs), _lib.ptr(rs), _lib.current_stream()),
               "mz_build_targets")
    torch.cuda.synchronize()
    assert np.array_equal(obs.cpu().numpy(), g["b%d_obs" % b])            # bit-exact gathers
    assert np.array_equal(acts.cpu().numpy(), g["b%d_actions" % b])
    assert np.array_equal(tr.cpu().numpy(), g["b%d_t_rewards" % b])
    assert np.array_equal(tp.cpu().numpy(), g["b%d_t_policies" % b])
    want = g["b%d_t_values" % b]
    err = np.max(np.abs(tv.cpu().numpy() - want) / np.maximum(np.abs(want), 1.0))
    worst = max(worst, err)
    assert err <= 1e-5  # north_star tolerance for n-step targets (np.dot order is BLAS-defined)
    if fuse:  # supports == value_phi(scalar_transform(targets)) on our own kernels (learners.py:186-192)
      v2 = cfgobj.value_phi(Config.scalar_transform(tv))
      r2 = cfgobj.reward_phi(Config.scalar_transform(tr))
      torch.cuda.synchronize()
      assert torch.equal(vs, v2) and torch.equal(rs, r2)
      assert torch.allclose(vs.sum(-1), torch.ones_like(vs.sum(-1)), atol=1e-6)
  print("worst n-step value rel err", worst)


def test_transform_kernels_match_reference():
  from model_based_rl_b200.config import Config
  g = load("transforms")
  cfg = Config(dict(value_support=[-15, 15], reward_support=[-15, 15], no_target_transform=False))
  h = Config.scalar_transform(torch.from_numpy(g["x"]).cuda()).cpu().numpy()
  # torch's CPU sqrt is not correctly rounded (1 ulp off on ~0.7% of inputs): few-ulp tolerance
  assert np.mean(h == g["h"]) > 0.98
  assert np.max(np.abs(h - g["h"]) / np.maximum(np.abs(g["h"]), 1e-3)) < 1e-6
  x = torch.from_numpy(g["support_in"].copy()).cuda().contiguous()
  sup = cfg.value_phi(x)
  torch.cuda.synchronize()
  assert np.array_equal(sup.cpu().numpy(), g["support"])                  # two-hot is bit exact
  assert np.array_equal(x.cpu().numpy(), np.clip(g["support_in"], -15, 15))  # x.clamp_ in place
  inv = cfg.inverse_value_transform(torch.from_numpy(g["logits"]).cuda()).cpu().numpy()
  rel = np.abs(inv - g["inverse"]) / np.maximum(np.abs(g["inverse"]), 1.0)
  # h^-1 in float32 quantises in ~1e-4 steps; 1-ulp differences in the expectation move one step
  assert np.mean(rel <= 1e-5) > 0.95 and rel.max() < 5e-4
  # which side is off?  Against the binary64 restatement of config.py:27-33 (oracle.inverse_transform_f64) the
  # kernel is no further from the exact value than the reference's own float32 torch pipeline
  truth = oracle.inverse_transform_f64(g["logits"], -15, 15)
  scale = np.maximum(np.abs(truth), 1.0)
  err_kernel, err_torch = np.abs(inv.reshape(-1) - truth) / scale, np.abs(g["inverse"].reshape(-1) - truth) / scale
  print("h^-1 vs binary64: kernel max %.3g mean %.3g | torch max %.3g mean %.3g" % (
      err_kernel.max(), err_kernel.mean(), err_torch.max(), err_torch.mean()))
  assert err_kernel.max() <= 1.25 * err_torch.max() + 1e-6
  assert err_kernel.mean() <= 1.25 * err_torch.mean() + 1e-7
  cfg_nt = Config(dict(value_support=[-15, 15], reward_support=[-15, 15], no_target_transform=True))
  inv_nt = cfg_nt.inverse_value_transform(torch.from_numpy(g["logits"]).cuda()).cpu().numpy()
  assert np.allclose(inv_nt, g["inverse_no_transform"], rtol=1e-5, atol=1e-6)


def test_full_size_target_properties():
  """Breakout-scale window (200k positions), B=512: linearity / structural invariants."""
  from model_based_rl_b200 import _lib
  lib = _lib.load()
  rng = np.random.default_rng(3)
  P, A, K, T, B, E = 200_000, 4, 5, 10, 512, 128
  lens = rng.integers(200, 800, size=P // 200)
  lens = lens[np.cumsum(lens) <= P]
  starts = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int64)
  obs = torch.from_numpy(rng.integers(0, 256, size=(P, E), dtype=np.uint8)).cuda()
  rewards = torch.from_numpy(np.sign(rng.normal(size=P) * (rng.random(P) < 0.1)).astype(np.float32)).cuda()
  actions = torch.from_numpy(rng.integers(0, A, size=P, dtype=np.int32)).cuda()
  to_play = torch.ones(P, dtype=torch.int8, device="cuda")
  root_values = torch.from_numpy(rng.normal(0, 2, size=P)).cuda()
  cv = rng.random((P, A)).astype(np.float32)
  cv /= cv.sum(1, keepdims=True)
  child_visits = torch.from_numpy(cv).cuda()
  win = _lib.Window(A, E, 1, 0, obs.data_ptr(), actions.data_ptr(), rewards.data_ptr(),
                    to_play.data_ptr(), root_values.data_ptr(), child_visits.data_ptr())
  disc = 0.997
  discounts = torch.from_numpy(np.array([disc**n for n in range(K + T)], np.float32)).cuda()
  ci = rng.integers(0, len(lens), size=B)
  step = (rng.random(B) * lens[ci]).astype(np.int64)
  pos = torch.from_numpy(starts[ci] + step).cuda()
  cs, cl = torch.from_numpy(starts[ci]).cuda(), torch.from_numpy(lens[ci].astype(np.int32)).cuda()
  pads = torch.from_numpy(rng.integers(0, A, size=(B, K), dtype=np.int32)).cuda()

  def run(rv):
    w = _lib.Window(A, E, 1, 0, obs.data_ptr(), actions.data_ptr(), rewards.data_ptr(),
                    to_play.data_ptr(), rv.data_ptr(), child_visits.data_ptr())
    tc = _lib.TargetCfg(B, K, T, 0, -15, 15, -15, 15, 0, 0, disc**T, discounts.data_ptr(), None, None)
    out = [torch.zeros((B, E), device="cuda"), torch.zeros((B, K), dtype=torch.int32, device="cuda"),
           torch.zeros((B, K + 1), device="cuda"), torch.zeros((B, K + 1), device="cuda"),
           torch.zeros((B, K + 1, A), device="cuda")]
    _lib.check(lib.mz_build_targets(w, tc, _lib.ptr(pos), _lib.ptr(cs), _lib.ptr(cl), _lib.ptr(pads),
                                    *[_lib.ptr(o) for o in out], None, None, _lib.current_stream()), "t")
    torch.cuda.synchronize()
    return out

  o1 = run(root_values)
  o0 = run(torch.zeros_like(root_values))
  # linearity in the bootstrap: value(rv) - value(0) == f32(rv[t+T] * disc^T) where it exists
  tpos = (pos.cpu().numpy()[:, None] + np.arange(K + 1)[None, :])
  in_chunk = (step[:, None] + np.arange(K + 1)[None, :] + T) < lens[ci][:, None]
  boot = np.where(in_chunk, root_values.cpu().numpy()[np.minimum(tpos + T, P - 1)] * disc**T, 0.0)
  diff = (o1[3] - o0[3]).cpu().numpy()
  assert np.allclose(diff, boot, rtol=1e-5, atol=1e-5)
  assert np.array_equal(o1[0].cpu().numpy(), obs.cpu().numpy()[pos.cpu().numpy()].astype(np.float32))
  beyond = (step[:, None] + np.arange(K + 1)[None, :]) >= lens[ci][:, None]
  assert (o1[4].cpu().numpy()[beyond] == 0).all() and (o1[3].cpu().numpy()[beyond] == 0).all()
  assert np.allclose(o1[4].cpu().numpy()[~beyond].sum(-1), 1.0, atol=1e-5)


@pytest.mark.parametrize("shape", [
    # A, K, T, B, E, obs_u8, normalise, two players
    (4, 5, 10, 13, 7, True, True, False),      # odd byte observations, batch not a multiple of the CTA's rows
    (9, 5, 10, 64, 9, False, False, True),     # Tic-Tac-Toe shape: sign flips by to_play
    (3, 0, 3, 5, 4, False, True, False),       # no unroll steps
    (2, 40, 1000, 9, 8, False, False, True),   # more unroll positions than lanes, LunarLander's td_steps
    (18, 5, 1, 33, 12, True, False, False),    # td_steps = 1
    (4, 15, 64, 7, 128, True, False, True),    # the lane-per-position kernel's limits: 16 positions, td_steps 64
    (6, 9, 7, 31, 16, False, True, True),      # three rows per warp, two idle lanes
    (4, 5, 10, 48, 16, True, True, True),      # TMA-staged kernel, every bulk path on: full CTAs, byte observations
    (8, 3, 6, 35, 8, False, False, False),     # same with float32 observations landing in the output image directly
    (4, 2, 5, 32, 64, False, True, False),     # float32 observations normalised in place in shared memory
])
def test_build_targets_ragged_shapes_match_oracle(shape, targets_kernel):
  """Every row of a batch against oracle.insert_target (replay_buffer.py:165-198) on short ragged chunks:
  positions at and past the chunk end, first position of a chunk, shapes off the vectorised paths."""
  import oracle
  from model_based_rl_b200 import _lib
  from model_based_rl_b200.config import Config
  A, K, T, B, E, u8, norm, two = shape
  lib = _lib.load()
  rng = np.random.default_rng(A * 1000 + K * 10 + T)
  lens = rng.integers(1, 60, size=40).astype(np.int32)
  lens[:3] = [1, 2, K + T + 5]
  starts = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int64)
  P = int(lens.sum())
  obs_np = rng.integers(0, 256, size=(P, E), dtype=np.uint8) if u8 else rng.normal(size=(P, E)).astype(np.float32)
  rewards = (rng.normal(size=P) * (rng.random(P) < 0.5)).astype(np.float32)
  to_play = (rng.choice([-1, 1], size=P) if two else np.ones(P)).astype(np.int8)
  root_values = rng.normal(0, 2, size=P)
  cv = rng.random((P, A)).astype(np.float32)
  cv /= cv.sum(1, keepdims=True)
  actions = rng.integers(0, A, size=P, dtype=np.int32)
  disc = 0.997
  ci = rng.integers(0, len(lens), size=B)
  ci[:3] = [0, 1, 2]
  step = (rng.random(B) * lens[ci]).astype(np.int64)
  step[2] = 0
  step[-1] = lens[ci[-1]] - 1
  pads_np = rng.integers(0, A, size=(B, max(K, 1)), dtype=np.int32)[:, :K].copy()
  omin = rng.normal(size=E).astype(np.float32)
  orng = (rng.random(E) + 0.5).astype(np.float32)
  d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
  t_obs, t_act, t_rew, t_tp, t_rv, t_cv = d(obs_np), d(actions), d(rewards), d(to_play), d(root_values), d(cv)
  t_min, t_rng = d(omin), d(orng)
  discounts = d(np.array([disc**n for n in range(K + T)], np.float32))
  win = _lib.Window(A, E, int(u8), 0, t_obs.data_ptr(), t_act.data_ptr(), t_rew.data_ptr(), t_tp.data_ptr(),
                    t_rv.data_ptr(), t_cv.data_ptr())
  tc = _lib.TargetCfg(B, K, T, 1, -15, 15, -15, 15, 0, int(norm), disc**T, discounts.data_ptr(),
                      t_min.data_ptr() if norm else None, t_rng.data_ptr() if norm else None)
  pos, cs, cl = d(starts[ci] + step), d(starts[ci]), d(lens[ci])
  pads = d(pads_np) if K else torch.zeros(1, dtype=torch.int32, device="cuda")
  out = [torch.full((B, E), -7.0, device="cuda"), torch.full((B, max(K, 1)), -7, dtype=torch.int32, device="cuda")[:, :K].contiguous(),
         torch.full((B, K + 1), -7.0, device="cuda"), torch.full((B, K + 1), -7.0, device="cuda"),
         torch.full((B, K + 1, A), -7.0, device="cuda"), torch.full((B, K + 1, 31), -7.0, device="cuda"),
         torch.full((B, K + 1, 31), -7.0, device="cuda")]
  ptrs = [_lib.ptr(o) if o.numel() else _lib.ptr(pads) for o in out]
  _lib.check(lib.mz_build_targets(win, tc, _lib.ptr(pos), _lib.ptr(cs), _lib.ptr(cl), _lib.ptr(pads), *ptrs,
                                  _lib.current_stream()), "mz_build_targets")
  torch.cuda.synchronize()
  obs, acts, tr, tv, tp, vs, rs = [o.cpu().numpy() for o in out]
  for b in range(B):
    lo, n, s = int(starts[ci[b]]), int(lens[ci[b]]), int(step[b])
    sl = slice(lo, lo + n)
    wr, wv, wp = oracle.insert_target(rewards[sl], to_play[sl], root_values[sl], cv[sl], K, T, disc, s)
    assert np.array_equal(tr[b], wr), (b, tr[b], wr)
    assert np.array_equal(tp[b], wp)
    assert np.max(np.abs(tv[b] - wv) / np.maximum(np.abs(wv), 1.0)) <= 1e-5
    want_obs = obs_np[lo + s].astype(np.float32)
    if norm:
      want_obs = (want_obs - omin) / orng
    assert np.array_equal(obs[b], want_obs)
    real = actions[lo + s:lo + min(s + K, n)]
    assert np.array_equal(acts[b], np.concatenate([real, pads_np[b, :K - len(real)]]))
  cfgobj = Config(dict(value_support=[-15, 15], reward_support=[-15, 15], no_target_transform=False))
  assert torch.equal(out[5], cfgobj.value_phi(Config.scalar_transform(out[3])))
  assert torch.equal(out[6], cfgobj.reward_phi(Config.scalar_transform(out[2])))


def test_bulk_launch_kernels_agree_and_match_oracle():
  """65 536 rows in one launch (128 learner batches, the shape bench.py's `targets.bulk` times) over a 200 000
  position window: the lane-per-position kernel and the warp-per-row kernel are independent implementations and
  must agree everywhere (gathers, reward / policy targets and supports bit for bit, n-step values to 1e-6
  relative -- they sum in different orders in float64); 300 random rows are checked against the oracle."""
  import oracle
  from model_based_rl_b200 import _lib
  lib = _lib.load()
  rng = np.random.default_rng(11)
  P, A, K, T, B, E = 200_000, 4, 5, 10, 65_536, 128
  lens = rng.integers(200, 800, size=P // 200)
  lens = lens[np.cumsum(lens) <= P].astype(np.int32)
  starts = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int64)
  obs_np = rng.integers(0, 256, size=(P, E), dtype=np.uint8)
  rewards = np.sign(rng.normal(size=P) * (rng.random(P) < 0.3)).astype(np.float32)
  to_play = rng.choice([-1, 1], size=P).astype(np.int8)
  root_values = rng.normal(0, 2, size=P)
  cv = rng.random((P, A)).astype(np.float32)
  cv /= cv.sum(1, keepdims=True)
  actions = rng.integers(0, A, size=P, dtype=np.int32)
  d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
  t_obs, t_act, t_rew, t_tp, t_rv, t_cv = d(obs_np), d(actions), d(rewards), d(to_play), d(root_values), d(cv)
  disc = 0.997
  discounts = d(np.array([disc**n for n in range(K + T)], np.float32))
  win = _lib.Window(A, E, 1, 0, t_obs.data_ptr(), t_act.data_ptr(), t_rew.data_ptr(), t_tp.data_ptr(),
                    t_rv.data_ptr(), t_cv.data_ptr())
  tc = _lib.TargetCfg(B, K, T, 1, -15, 15, -15, 15, 0, 0, disc**T, discounts.data_ptr(), None, None)
  ci = rng.integers(0, len(lens), size=B)
  step = (rng.random(B) * lens[ci]).astype(np.int64)
  step[:64] = lens[ci[:64]] - 1 - (np.arange(64) % 8)          # rows that run past the end of their chunk
  pos, cs, cl = d(starts[ci] + step), d(starts[ci]), d(lens[ci])
  pads = d(rng.integers(0, A, size=(B, K), dtype=np.int32))

  def run(which):
    lib.mz_debug_set_targets_kernel(which)
    out = [torch.full((B, E), -7.0, device="cuda"), torch.full((B, K), -7, dtype=torch.int32, device="cuda"),
           torch.full((B, K + 1), -7.0, device="cuda"), torch.full((B, K + 1), -7.0, device="cuda"),
           torch.full((B, K + 1, A), -7.0, device="cuda"), torch.full((B, K + 1, 31), -7.0, device="cuda"),
           torch.full((B, K + 1, 31), -7.0, device="cuda")]
    try:
      _lib.check(lib.mz_build_targets(win, tc, _lib.ptr(pos), _lib.ptr(cs), _lib.ptr(cl), _lib.ptr(pads),
                                      *[_lib.ptr(o) for o in out], _lib.current_stream()), "mz_build_targets")
      torch.cuda.synchronize()
    finally:
      lib.mz_debug_set_targets_kernel(0)
    return out
  rows, wide, lanes = run(0), run(1), run(2)
  # the TMA-staged kernel (by shape) and the lane-per-position kernel do the same arithmetic in the same order
  for i in range(7):
    assert torch.equal(rows[i], lanes[i]), i
  for i in (0, 1, 2, 4):
    assert torch.equal(rows[i], wide[i]), i
  rel = ((rows[3] - wide[3]).abs() / wide[3].abs().clamp(min=1.0)).max().item()
  assert rel <= 1e-6, rel
  same = rows[3] == wide[3]
  assert torch.equal(rows[5][same], wide[5][same]) and torch.equal(rows[6], wide[6])
  assert torch.allclose(rows[5].sum(-1), torch.ones(B, K + 1, device="cuda"), atol=1e-6)
  assert torch.equal(rows[0], t_obs[pos].float())
  tr, tv, tp = rows[2].cpu().numpy(), rows[3].cpu().numpy(), rows[4].cpu().numpy()
  for b in list(range(64)) + rng.integers(0, B, size=236).tolist():
    lo, n, s = int(starts[ci[b]]), int(lens[ci[b]]), int(step[b])
    sl = slice(lo, lo + n)
    wr, wv, wp = oracle.insert_target(rewards[sl], to_play[sl], root_values[sl], cv[sl], K, T, disc, s)
    assert np.array_equal(tr[b], wr) and np.array_equal(tp[b], wp)
    assert np.max(np.abs(tv[b] - wv) / np.maximum(np.abs(wv), 1.0)) <= 1e-5
