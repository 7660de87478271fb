"""GPU parity tests for the target kernel and the scalar transform kernels against goldens produced
by the unmodified reference (replay_buffer.PrioritizedReplay.sample_batch, config.Config)."""
import ctypes as C
import types

import numpy as np
import pytest
import torch

from helpers import REPLAY_CASES, load

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


@pytest.mark.parametrize("case", REPLAY_CASES)
@pytest.mark.parametrize("fuse", [False, True])
def test_build_targets_matches_reference(case, fuse):
  from model_based_rl_b200 import _lib
  from model_based_rl_b200.config import Config
  lib = _lib.load()
  g = load("replay_" + case)
  A, K, T, B = int(g["action_space"]), int(g["num_unroll_steps"]), int(g["td_steps"]), int(g["batch_size"])
  E, disc = int(g["obs_dim"]), float(g["discount"])
  dev, starts, lens = _window_from_golden(g)
  win = _lib.Window(A, E, int(g["obs_uint8"]), 0, dev["obs"].data_ptr(), dev["actions"].data_ptr(),
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
