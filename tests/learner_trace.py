"""Diagnostics (not a test): per-phase clock64 stamps of CTA 0 of the chain_fwd kernel.  Needs a library built with
-DMZ_TC_TRACE:  MZB200_LIB=ab/lib_trace.so python tests/learner_trace.py"""
import ctypes as C, os, sys, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from model_based_rl_b200 import fused_learner, _lib
dev = torch.device("cuda:0"); B, K, A, E = 512, 5, 4, 128
cfg = types.SimpleNamespace(value_support=[-15, 15], reward_support=[-15, 15], no_support=False, no_target_transform=False,
                            num_unroll_steps=K, optimizer="AdamW", lr_init=0.0008, momentum=0.9, weight_decay=1e-4, clip_grad=0, lr_scheduler=None, norm_obs=False)
g = torch.Generator(device=dev).manual_seed(7); r = lambda *s: torch.rand(*s, device=dev, generator=g); pol = r(B, K + 1, A)
batch = ((r(B, E), torch.randint(0, A, (B, K), device=dev, generator=g), ((r(B, K + 1) < 0.1).float(), 4 * torch.randn(B, K + 1, device=dev, generator=g), pol / pol.sum(-1, keepdim=True))), None, r(B).double())
lr = fused_learner.FusedLearner(cfg, fused_learner.FusedFCNetwork(E, A, dev, cfg), use_graph=False)
for _ in range(6): lr.update_weights(batch)
torch.cuda.synchronize()
lib = _lib.load()
buf = np.zeros((2, 2, 16, 8), dtype=np.int64)
lib.mz_debug_tc_trace.restype = C.c_int
assert lib.mz_debug_tc_trace(buf.ctypes.data_as(C.c_void_p)) == 0
names = ["start", "hidden done", "sync1", "output done", "sync2", "LN done", "sync3"]
for wi, wn in enumerate(("warp 0", "last warp")):
  print(wn)
  for step in range(K + 1):
    t = buf[0, wi, step]
    print("  step %d: " % step + ", ".join("%s +%d" % (names[i], t[i] - t[i - 1]) for i in range(1, 7)) + "  | total %d" % (t[6] - t[0]))
print("whole chain (warp 0): %d cycles" % (buf[0, 0, K, 6] - buf[0, 0, 0, 0]))
hn = ["loads + gb2 + sync", "hidden + dH", "sync", "dX", "gW2", "gW1"]
for wi, wn in enumerate(("warp 0", "last warp")):
  t = buf[1, wi, 0]
  print("heads_bwd CTA (0, y) last launch, %s: " % wn + ", ".join("%s +%d" % (hn[i - 1], t[i] - t[i - 1]) for i in range(1, 7)) + " | total %d" % (t[6] - t[0]))
