"""Diagnostics (not a test): steps/s of FusedLearner against learners.Learner on the C3 shape (B = 512, K = 5)."""
import os, sys, time, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from model_based_rl_b200 import fused_learner, learners
dev = torch.device("cuda:0")
B, K, A, E = 512, 5, 4, 128
cfg = types.SimpleNamespace(value_support=[-15, 15], reward_support=[-15, 15], no_support=False, no_target_transform=False,
                            num_unroll_steps=K, optimizer=(sys.argv[1] if len(sys.argv) > 1 else "AdamW"), lr_init=0.0008,
                            momentum=0.9, weight_decay=1e-4, clip_grad=0, lr_scheduler=None, norm_obs=False)
g = torch.Generator(device=dev).manual_seed(7)
r = lambda *shape: torch.rand(*shape, device=dev, generator=g)
pol = r(B, K + 1, A)
batch = ((r(B, E), torch.randint(0, A, (B, K), device=dev, generator=g),
          ((r(B, K + 1) < 0.1).float(), 4 * torch.randn(B, K + 1, device=dev, generator=g), pol / pol.sum(-1, keepdim=True))),
         None, r(B).double())
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for name in ("bf16_graph", "bf16_eager", "f32_graph", "torch_graph"):
  if not name.startswith("torch"):
    lr = fused_learner.FusedLearner(cfg, fused_learner.FusedFCNetwork(E, A, dev, cfg), use_graph=name.endswith("graph"),
                                    precision=name.split("_")[0])
  else:
    lr = learners.Learner(cfg, learners.FCNetworkTrain(E, A, dev, cfg), use_graph=True)
  for _ in range(5):
    lr.update_weights(batch)
  torch.cuda.synchronize()
  t0 = time.perf_counter()
  a.record()
  n = 50
  for _ in range(n):
    lr.update_weights(batch)
  b.record()
  torch.cuda.synchronize()
  print("%-12s %8.1f us/step device, %8.1f us/step wall, %7.0f steps/s, loss %s" %
        (name, a.elapsed_time(b) * 1e3 / n, (time.perf_counter() - t0) * 1e6 / n, n / (a.elapsed_time(b) * 1e-3),
         lr.last_losses.cpu().numpy().round(4)))
