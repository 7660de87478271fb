"""Diagnostics (not a test): launch shapes of the TMA-staged target kernel against the lane-per-position kernel,
C3 shape (K=5, td=10, A=4, 128-byte observations, supports fused), 65 536 rows and one learner batch per launch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from model_based_rl_b200 import _lib
lib = _lib.load()
dev = torch.device("cuda:0")
P, A, K, T, E = 200_000, 4, 5, 10, 128
def case(B, reps):
  rng = np.random.default_rng(5)
  sec, bps = bench._targets_case(torch, _lib, dev, rng, P, A, K, T, B, E, True, 4, reps)
  return sec * 1e6, B * bps / sec / 1e9
for B, reps in ((65536, 10), (512, 50)):
  lib.mz_debug_set_targets_kernel(2)
  us, gbs = case(B, reps)
  print("B=%6d lane-per-position           %8.2f us  %7.1f GB/s  %.3f" % (B, us, gbs, gbs / 6545.9))
  lib.mz_debug_set_targets_kernel(0)
  for rows in (4, 8, 16, 32):
    for thr in (64, 128, 256):
      assert lib.mz_debug_set_targets_tma(rows, thr) == 0
      us, gbs = case(B, reps)
      print("B=%6d tma rows=%2d threads=%3d       %8.2f us  %7.1f GB/s  %.3f" % (B, rows, thr, us, gbs, gbs / 6545.9))
