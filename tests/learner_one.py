"""Diagnostics (not a test): the learner step alone (bench.py's `learner` object), e.g. under ncu / torch profiler."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from model_based_rl_b200 import _lib
print(json.dumps(bench.bench_learner(torch, _lib, torch.device("cuda:0"), 1, lambda: None), indent=1))
