# Launch list of FusedLearner steps (one B200): bash tests/learner_launches.sh r02z
R=${1:-r02z}
mkdir -p gpurun_out
cat > /tmp/ll.py <<'P'
import sys, types, torch
sys.path.insert(0, '.')
from model_based_rl_b200 import fused_learner
dev = torch.device("cuda:0"); B, K, A, E = 512, 5, 4, 128
cfg = types.SimpleNamespace(value_support=[-15, 15], reward_support=[-15, 15], no_support=False, no_target_transform=False,
                            num_unroll_steps=K, optimizer="AdamW", lr_init=0.0008, momentum=0.9, weight_decay=1e-4, clip_grad=0, lr_scheduler=None, norm_obs=False)
g = torch.Generator(device=dev).manual_seed(7); r = lambda *s: torch.rand(*s, device=dev, generator=g); pol = r(B, K + 1, A)
batch = ((r(B, E), torch.randint(0, A, (B, K), device=dev, generator=g), ((r(B, K + 1) < 0.1).float(), 4 * torch.randn(B, K + 1, device=dev, generator=g), pol / pol.sum(-1, keepdim=True))), None, r(B).double())
lr = fused_learner.FusedLearner(cfg, fused_learner.FusedFCNetwork(E, A, dev, cfg), use_graph=False)
for _ in range(8): lr.update_weights(batch)
torch.cuda.synchronize()
P
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_learner_launches.csv python /tmp/ll.py > /dev/null 2>&1
python - <<P
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/${R}_learner_launches.csv")) if len(r) > 10]
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); gi = hdr.index("Grid Size")
d = collections.OrderedDict()
for r in rows[1:]:
  d.setdefault((r[ki][-40:], r[gi]), []).append(float(r[vi].replace(",", "")))
for k, v in d.items():
  print(k, len(v), "avg %.1f us" % (sum(v[len(v)//2:]) / len(v[len(v)//2:]) / 1e3))
P
