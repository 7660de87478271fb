"""Diagnostics (not a test): MinMax epochs per game and per-simulation tree-step times."""
import sys, os, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from model_based_rl_b200.networks import FCNetwork, FCSearch, random_state_dict
cfg = types.SimpleNamespace(num_simulations=50, action_space=18, two_players=False, discount=0.997, pb_c_base=19652,
                            pb_c_init=1.25, init_value_score=0.0, known_bounds=[None, None], root_exploration_fraction=0.25,
                            value_support=[-15, 15], reward_support=[-15, 15], no_support=False, no_target_transform=False)
G, A, S = (int(sys.argv[1]) if len(sys.argv) > 1 else 4096), 18, 50
net = FCNetwork(128, A, "cuda", cfg)
net.load_weights(random_state_dict(128, A))
fs = FCSearch(cfg, net, G, use_graph=False, num_streams=1)
fs.eng.enable_trace()
rng = np.random.default_rng(0)
obs = (rng.integers(0, 256, size=(G, 128)).astype(np.float32) / 255.0); noise = rng.dirichlet([0.25] * A, size=G)
fs.search_host(obs, noise, rng.random(G), np.ones(G))
eng = fs.eng
games = eng.games.view(torch.uint8).reshape(G, -1)
ln = fs.lanes[0]
plan = ln._plan(fs.use_noise, fs.noise_frac, torch.cuda.current_stream().cuda_stream)
for fn, args in plan: fn(*args)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(plan) + 1)]
ev[0].record()
eps = []
for i, (fn, args) in enumerate(plan):
  fn(*args); ev[i + 1].record()
torch.cuda.synchronize()
dur = [ev[i].elapsed_time(ev[i + 1]) * 1e3 for i in range(len(plan))]
print("tree step us by sim:", [round(dur[4 + 2 * s], 1) for s in range(0, S, 3)])
print("fc us by sim:", [round(dur[3 + 2 * s], 1) for s in range(0, S, 6)])
print("path_len mean", float(eng.path_len.float().mean()))

# per-simulation step time against the mean depth of that simulation's descent: time = a + b * depth
depth = eng.trace[2].float().mean(dim=1).cpu().numpy()  # [S + 1] descents (the first one precedes simulation 0)
t = np.array([dur[4 + 2 * s] for s in range(S - 1)])   # step s = backup of simulation s + descent of simulation s + 1
d = depth[1:S]
b, a = np.polyfit(d, t, 1)
print("games %d: tree step = %.2f us + %.3f us per level (mean depth %.1f -> %.1f us); all sims:" % (G, a, b, d.mean(), t.mean()))
print(" ".join("%.1f/%.0f" % (x, y) for x, y in zip(d, t)))
