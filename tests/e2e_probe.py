"""Diagnostics (not a test): variants of the one-graph end-to-end call of FCSearch (search_pinned)."""
import os, sys, time, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from model_based_rl_b200 import _lib
from model_based_rl_b200.networks import FCNetwork, FCSearch, random_state_dict
args = types.SimpleNamespace(games=4096, sims=50, actions=18, obs_dim=128, precision="bf16")
dev = torch.device("cuda:0")
cfg = bench.search_config(args)
net = FCNetwork(args.obs_dim, args.actions, dev, cfg, precision="bf16")
net.load_weights({k: v.to(dev) for k, v in random_state_dict(args.obs_dim, args.actions, seed=1234).items()})
fs = FCSearch(cfg, net, args.games)
obs, noise, uniforms, temperature = bench.synthetic_inputs(args, 0, args.games)
obs_u8 = bench.synthetic_inputs(args, 0, args.games, as_bytes=True)[0]
pinned = fs.pinned_inputs()
for name, src in (("obs_u8", obs_u8), ("noise", noise), ("uniforms", uniforms), ("temperature", temperature)):
  pinned[name].copy_(torch.from_numpy(src))
fs.use_noise = True
mn, rg = getattr(fs, '_obs_norm', (None, None))
split = fs.obs_u8.data_ptr() - fs._in_dev.data_ptr()
side = torch.cuda.Stream()

def norm():
  _lib.check(net.lib.mz_obs_normalize_u8(fs.G, net.input_dim, _lib.ptr(fs.obs_u8), _lib.ptr(mn), _lib.ptr(rg), _lib.ptr(fs.obs),
                                         _lib.current_stream()), "norm")

def variant(h2d, fork, d2h):
  def body():
    main = torch.cuda.current_stream()
    if h2d and fork:
      side.wait_stream(main)
      with torch.cuda.stream(side):
        fs._in_dev[:split].copy_(fs._in_host[:split], non_blocking=True)
      fs._in_dev[split:].copy_(fs._in_host[split:], non_blocking=True)
    elif h2d:
      fs._in_dev.copy_(fs._in_host, non_blocking=True)
    norm()
    fs._enqueue()
    if d2h:
      fs._out_host.copy_(fs._out_dev, non_blocking=True)
  body(); torch.cuda.synchronize()
  g = torch.cuda.CUDAGraph()
  with torch.cuda.graph(g):
    body()
  def call():
    if not h2d:
      fs._in_dev.copy_(fs._in_host, non_blocking=True)
    g.replay()
    if not d2h:
      fs._out_host.copy_(fs._out_dev, non_blocking=True)
    torch.cuda.current_stream().synchronize()
  return call

def old():
  fs._in_dev.copy_(fs._in_host, non_blocking=True)
  norm()
  fs.run()
  fs._out_host.copy_(fs._out_dev, non_blocking=True)
  torch.cuda.current_stream().synchronize()

def timeit(name, fn, n=50):
  for _ in range(5): fn()
  torch.cuda.synchronize(); t0 = time.perf_counter()
  for _ in range(n): fn()
  torch.cuda.synchronize()
  print("%-44s %8.1f us per call" % (name, (time.perf_counter() - t0) / n * 1e6))

timeit("run() only (inputs resident)", lambda: (fs.run(), torch.cuda.synchronize()))
timeit("old search_pinned", old)
timeit("graph: norm + move (copies outside)", variant(False, False, False))
timeit("graph: + H2D (one copy)", variant(True, False, False))
timeit("graph: + H2D (one copy) + D2H", variant(True, False, True))
