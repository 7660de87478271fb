"""Diagnostics (not a test): per-launch times of the conv path at bench scale."""
import sys, os, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from model_based_rl_b200 import _lib
from model_based_rl_b200.muzero import MuZeroNetwork, random_state_dict, ROWS, CH
cfg = types.SimpleNamespace(value_support=[-15, 15], reward_support=[-15, 15], no_support=False, no_target_transform=False)
G = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
_lib.load().mz_conv_set_pair(int(os.environ.get("MZ_CONV_PAIR", "2")))  # 0: single-CTA kernel (A/B runs)
net = MuZeroNetwork(32, 18, "cuda", cfg)
net.load_weights(random_state_dict(32, 18))
x = torch.rand((G * ROWS, CH), device="cuda").to(torch.bfloat16)
nxt = torch.zeros_like(x)
acts = torch.randint(0, 18, (G,), device="cuda", dtype=torch.int32)
v = torch.zeros(G, device="cuda"); r = torch.zeros(G, device="cuda"); l = torch.zeros((G, 18), device="cuda")
def ev(fn, n=5):
  for _ in range(2): fn()
  torch.cuda.synchronize()
  a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  a.record()
  for _ in range(n): fn()
  b.record(); torch.cuda.synchronize()
  return a.elapsed_time(b) / n * 1e3
bufs = net.buffers(G)["x"]
conv = net.dyn_tower[0]
t_conv = ev(lambda: net._conv(G, conv, x, 1, bufs[0]), 20)
t_res = ev(lambda: net._conv(G, conv, x, 3, bufs[0], residual=x), 20)
t_rec = ev(lambda: net.run_recurrent(G, x, acts, v, r, l))
flops_conv = 2.0 * G * 36 * 128 * 1152
flops_issued = 2.0 * G * ROWS * 128 * 1152
print("G=%d conv3x3: %.1f us (%.0f TFLOP/s useful, %.0f issued), with residual %.1f us" % (G, t_conv, flops_conv / t_conv / 1e6, flops_issued / t_conv / 1e6, t_res))
print("recurrent_inference: %.1f us total = %.2f us/game ; useful %.0f TFLOP/s" % (t_rec, t_rec / G, G * 0.705e9 / t_rec / 1e6))

