"""Diagnostics (not a test): the replay facade section of bench.py alone."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from model_based_rl_b200 import _lib
print(json.dumps(bench.bench_replay(torch, _lib, torch.device("cuda:0"), cpu_baseline=False), indent=1))
