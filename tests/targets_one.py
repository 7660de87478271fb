"""Diagnostics (not a test): one configuration of the target kernels for ncu.  argv: which rows threads B [c2]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from model_based_rl_b200 import _lib
lib = _lib.load()
which, rows, thr, B = [int(x) for x in sys.argv[1:5]]
c2 = len(sys.argv) > 5 and sys.argv[5] == "c2"
lib.mz_debug_set_targets_kernel(which)
assert lib.mz_debug_set_targets_tma(rows, thr) == 0
if c2:  # LunarLander: td_steps = 1000, 8 float32 observations
  sec, bps = bench._targets_case(torch, _lib, torch.device("cuda:0"), np.random.default_rng(5), 200_000, 4, 5, 1000, B, 8, False, 4, 2)
else:
  sec, bps = bench._targets_case(torch, _lib, torch.device("cuda:0"), np.random.default_rng(5), 200_000, 4, 5, 10, B, 128, True, 4, 2)
print(sec * 1e6, "us")
