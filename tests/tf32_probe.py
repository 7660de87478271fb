"""Diagnostics: microseconds per recurrent_inference launch of the three FCNetwork kernels (CUDA-core float32,
split-TF32 tensor cores, bf16 tcgen05) at the batch sizes of the search, and the whole move with each.
   python tests/tf32_probe.py [A] [obs_dim]"""
import sys, types
import numpy as np, torch
sys.path.insert(0, ".")
from model_based_rl_b200.networks import FCNetwork, FCSearch, random_state_dict

A = int(sys.argv[1]) if len(sys.argv) > 1 else 18
D = int(sys.argv[2]) if len(sys.argv) > 2 else 128
cfg = types.SimpleNamespace(
    num_simulations=50, action_space=A, two_players=False, discount=0.997, pb_c_base=19652, pb_c_init=1.25,
    init_value_score=0.0, known_bounds=[None, None], root_dirichlet_alpha=0.25, root_exploration_fraction=0.25,
    value_support=[-15, 15], reward_support=[-15, 15], no_support=False, no_target_transform=False)
sd = {k: v.cuda() for k, v in random_state_dict(D, A, seed=1234).items()}
nets = {}
for prec in ("f32", "tf32x3", "bf16"):
  nets[prec] = FCNetwork(D, A, "cuda", cfg, precision=prec)
  nets[prec].load_weights(sd)

def timed(fn, n=50):
  for _ in range(5):
    fn()
  torch.cuda.synchronize()
  a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  a.record()
  for _ in range(n):
    fn()
  b.record()
  torch.cuda.synchronize()
  return a.elapsed_time(b) / n * 1e3

for B in (1024, 4096, 16384):
  h = torch.rand((B, 50), device="cuda")
  act = torch.randint(0, A, (B,), device="cuda", dtype=torch.int32)
  obs = torch.rand((B, D), device="cuda")
  line = []
  for prec, net in nets.items():
    v, r = torch.empty((B, 1), device="cuda"), torch.empty((B, 1), device="cuda")
    l, ho = torch.empty((B, A), device="cuda"), torch.empty((B, 50), device="cuda")
    us = timed(lambda: net.recurrent_into(h, 50, None, act, ho, 50, 0, v, r, l))
    ui = timed(lambda: net.initial_inference(obs))
    line.append("%s rec %.1f us init %.1f us" % (prec, us, ui))
  print("B=%d: " % B + " | ".join(line), flush=True)

G, S = 4096, 50
rng = np.random.default_rng(1)
for prec in ("f32", "tf32x3"):
  fs = FCSearch(cfg, nets[prec], G, use_graph=True, num_streams=4)
  fs.search_host(rng.integers(0, 256, (G, D)).astype(np.uint8), rng.dirichlet([0.25] * A, size=G),
                 rng.random(G), np.ones(G))
  us = timed(fs.run, 10)
  print("%s move: %.3f ms = %.1f M expansions/s" % (prec, us / 1e3, G * S / us), flush=True)
