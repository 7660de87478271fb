"""Diagnostics (not a test): the target kernel alone (bench.py's `targets` object), e.g. under ncu."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from model_based_rl_b200 import _lib
print(json.dumps(bench.bench_targets(torch, _lib, torch.device("cuda:0")), indent=1))
