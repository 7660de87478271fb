"""The device's exp (priors of Node.expand, mcts.py:52) must equal the host libm's exp -- what
math.exp calls -- bit for bit on float32-valued logits."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_device_exp_is_bit_identical_to_libm():
  from model_based_rl_b200 import _lib
  lib = _lib.load()
  rng = np.random.default_rng(11)
  n = 4_000_000
  bits = rng.integers(0, 2**32, size=n, dtype=np.uint64).astype(np.uint32)
  x = bits.view(np.float32)
  x = x[np.isfinite(x) & (np.abs(x) < 512)]
  x = np.concatenate([x, rng.normal(0, 3, size=2_000_000).astype(np.float32),
                      np.array([0.0, -0.0, 1.0, -1.0, 88.0, -88.0, 1e-30, 511.9], np.float32)])
  xd = torch.from_numpy(x).cuda()
  out = torch.empty(len(x), dtype=torch.float64, device="cuda")
  _lib.check(lib.mz_exp_f32(len(x), _lib.ptr(xd), _lib.ptr(out), _lib.current_stream()), "mz_exp_f32")
  torch.cuda.synchronize()
  want = np.exp(x.astype(np.float64))  # numpy float64 exp may be SIMD: compare against math.exp too
  got = out.cpu().numpy()
  import math
  idx = rng.integers(0, len(x), size=200_000)
  ref = np.array([math.exp(float(v)) for v in x[idx]])
  assert np.array_equal(got[idx], ref), "device exp differs from math.exp"
  print("mismatches vs np.exp (may be a SIMD variant): %d of %d" % (int((got != want).sum()), len(x)))
