"""Where a DeviceActor move spends its host time (C4 shape): cProfile over 30 moves."""
import cProfile
import os
import pstats
import sys
import time
import types

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from model_based_rl_b200.environments import SyntheticRam
from model_based_rl_b200.networks import FCNetwork, FCSearch, random_state_dict
from model_based_rl_b200.replay_buffer import PrioritizedReplay
from model_based_rl_b200.selfplay import DeviceActor

G, A, S, D = 4096, 18, 50, 128
cfg = types.SimpleNamespace(
    num_simulations=S, action_space=A, two_players=False, discount=0.997, pb_c_base=19652, pb_c_init=1.25,
    init_value_score=0.0, known_bounds=[None, None], root_dirichlet_alpha=0.25, root_exploration_fraction=0.25,
    value_support=[-15, 15], reward_support=[-15, 15], no_support=False, no_target_transform=False,
    num_unroll_steps=5, td_steps=10, max_history_length=500, max_steps=27000, batch_size=512,
    beta_increment_per_sampling=0.001, epsilon=0.01, alpha=1.0, beta=0.4, obs_space=(D,), window_size=200_000,
    window_step=None, seed=None, clip_rewards=True)
dev = "cuda:0"
net = FCNetwork(D, A, dev, cfg)
net.load_weights({k: v.to(dev) for k, v in random_state_dict(D, A, seed=5).items()})
env = SyntheticRam(G, A, D, episode_length=600, seed=1)
rb = PrioritizedReplay(cfg, device=dev, window_positions=int(200_000 * 1.3) + 3 * G * 515)
fs = FCSearch(cfg, net, G)
actor = DeviceActor(cfg, env, rb, fs)
env.elapsed[:] = np.arange(G) % 600
for _ in range(3):
  actor.play_move()
torch.cuda.synchronize()
moves = int(sys.argv[1]) if len(sys.argv) > 1 else 30
t0 = time.perf_counter()
for _ in range(moves):
  actor.play_move()
torch.cuda.synchronize()
print("ms per move: %.3f" % ((time.perf_counter() - t0) / moves * 1e3))
pr = cProfile.Profile()
pr.enable()
for _ in range(moves):
  actor.play_move()
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(28)
st.sort_stats("cumulative").print_stats(28)

# ---- coarse wall-clock sections (Cython / ctypes calls are invisible to cProfile) ----
def timeit(name, fn, n=20):
  fn()
  torch.cuda.synchronize()
  t0 = time.perf_counter()
  for _ in range(n):
    fn()
  torch.cuda.synchronize()
  print("%-34s %8.3f ms" % (name, (time.perf_counter() - t0) / n * 1e3))

legal = env.legal_mask()
timeit("dirichlet [G, A]", lambda: np.random.dirichlet([0.25] * A, size=G))
timeit("uniforms", lambda: np.random.random(G))
noise = np.zeros((G, A)); noise[:, :A] = np.random.dirichlet([0.25] * A, size=G)
u = np.random.random(G)
obs = actor.obs
timeit("search_host (u8 obs, noise)", lambda: fs.search_host(obs, noise, u, actor.temperature, legal=legal, to_play=actor.to_play))
timeit("search_pinned", lambda: fs.search_pinned())
timeit("fs.run only", lambda: fs.run())
acts = np.zeros(G, np.int32)
timeit("env.step", lambda: env.step(acts))
timeit("env._frames", lambda: env._frames(G))
def append():
  actor.h_pos.copy_(torch.from_numpy(actor.chunk_start))
  actor.d_pos.copy_(actor.h_pos, non_blocking=True)
  actor.d_rew.copy_(actor.h_rew, non_blocking=True)
  rb.append_steps(actor.d_pos, fs.obs_u8, fs.actions, actor.d_rew, fs.to_play, fs.root_value, fs.child_visits)
timeit("append_steps + staging", append)
errs = np.abs(np.random.normal(size=515))
def commit():
  cid, start = rb.open_chunk(515, torch.uint8)
  rb.commit_chunk(cid, start, 500, errs, ignore=15, terminal=False)
timeit("open_chunk + commit_chunk (500)", commit, 50)
timeit("copy_positions (8 runs)", lambda: rb.copy_positions([0] * 8, [1000 * i + 5000 for i in range(8)], [15] * 8))
