"""Diagnostics (not a test): prints the phase timeline of CTA 0 of the tensor-core network kernel."""
import sys, os, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from model_based_rl_b200 import _lib
from model_based_rl_b200.networks import FCNetwork, random_state_dict
cfg = types.SimpleNamespace(value_support=[-15, 15], reward_support=[-15, 15], no_support=False, no_target_transform=False)
A, B = 18, 4096
net = FCNetwork(128, A, "cuda", cfg, precision="bf16")
net.load_weights(random_state_dict(128, A))
lib = _lib.load()
trace = torch.zeros(512, dtype=torch.int64, device="cuda")
h = torch.rand(B, 50, device="cuda"); a = torch.randint(0, A, (B,), device="cuda", dtype=torch.int32)
for _ in range(3): net.recurrent_inference(h, a)
BLOCK = int(sys.argv[1]) if len(sys.argv) > 1 else 0
lib.mz_debug_set_tc_trace_block(BLOCK)
lib.mz_debug_set_tc_trace(_lib.ptr(trace))
net.recurrent_inference(h, a); torch.cuda.synchronize()
lib.mz_debug_set_tc_trace(None)
print("trace of CTA", BLOCK)
t = trace.cpu().numpy(); t0 = t[0]
rel = lambda i: (t[i] - t0) if t[i] else -1
print("epilogue: A1 ready", rel(1))
for c in range(16):
  m = [rel(64 + 8 * c + i) for i in range(7)]
  print("  chunk %2d: epi %6d..%6d | mma1: top %6d waits+%5d issue+%5d commit+%5d | mma2: start %6d issue+%5d commits+%5d | producer %6d" % (
      c, rel(4 + 2 * c), rel(5 + 2 * c), m[0], m[1] - m[0], m[2] - m[1], m[3] - m[2], m[4], m[5] - m[4], m[6] - m[5], rel(192 + c)))
print("d2_full(dyn)", rel(40), "a3 ready", rel(41), "d2_full(pred)", rel(42), "epi end", rel(43), "exit", rel(44))
print("A1 gather done per warp (2..9):", [int(rel(224 + w)) for w in range(2, 10)], "after fence:", [int(rel(236 + w)) for w in range(2, 10)])

# ---- the same trace inside a real move (hidden pool gather / scatter, L2 shared with the tree) ----
from model_based_rl_b200.networks import FCSearch
scfg = types.SimpleNamespace(num_simulations=50, action_space=A, two_players=False, discount=0.997, pb_c_base=19652,
                             pb_c_init=1.25, init_value_score=0.0, known_bounds=[None, None], root_exploration_fraction=0.25)
fs = FCSearch(scfg, net, 4096, use_graph=False, num_streams=1)
rng = np.random.default_rng(0)
obs = rng.random((4096, 128)).astype(np.float32); noise = rng.dirichlet([0.25] * A, size=4096)
fs.search_host(obs, noise, rng.random(4096), np.ones(4096))
trace.zero_()
lib.mz_debug_set_tc_trace(_lib.ptr(trace))
fs.search_host(obs, noise, rng.random(4096), np.ones(4096))
lib.mz_debug_set_tc_trace(None)
t = trace.cpu().numpy(); t0 = t[0]
print("IN-SEARCH (last simulation of a move):")
print("A1 gather done per warp (2..9):", [int(rel(224 + w)) for w in range(2, 10)])
for c in (0, 1, 7, 8, 9, 15):
  m = [rel(64 + 8 * c + i) for i in range(7)]
  print("  chunk %2d: epi %6d..%6d | mma1: top %6d waits+%5d issue+%5d | mma2: start %6d issue+%5d | producer %6d" % (
      c, rel(4 + 2 * c), rel(5 + 2 * c), m[0], m[1] - m[0], m[2] - m[1], m[4], m[5] - m[4], rel(192 + c)))
print("d2_full(dyn)", rel(40), "a3 ready", rel(41), "d2_full(pred)", rel(42), "epi end", rel(43), "exit", rel(44))
ent = t[256:256 + 64:2]; ext = t[257:257 + 64:2]
base = ent[ent > 0].min()
print("per-CTA entry (ns after first):", (ent[:32] - base).tolist())
print("per-CTA exit  (ns after first entry):", (ext[:32] - base).tolist())
